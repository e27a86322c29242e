"""pgasr_b200 -- the sequence-level training hot path of Policy-Gradient-ASR on B200 (sm_100a).

Host-side mirror of the upstream interface for this path (same module and function names, argument
meaning and error behaviour), backed by the CUDA kernels in csrc/ through the C ABI in include/pgasr.h:

    metrics.edit_dist / evaluate / save_predictions          upstream metrics.py
    CTCdecoder.collapse_fn / CTCDecoder                       upstream CTCdecoder.py
    policy_grad.reward                                        upstream policy_grad.py
    loss.customNLLLoss  (+ the new loss.PolicyGradCTCLoss)    upstream loss.py
    functional.*                                              batched tensor-level operators (device tensors)
    host.HostPipeline                                         the step on pinned HOST arrays, pipelined
    predict.decode_and_score                                  upstream predict()'s inner loop, batched (model.py:321-334)
    distributed.*                                             utterance sharding over ranks

There is no CPU fallback anywhere in this package.
"""
from . import _native                                        # noqa: F401
from . import functional                                     # noqa: F401
from . import metrics, CTCdecoder, policy_grad, loss, distributed, host, predict   # noqa: F401
from .loss import PolicyGradCTCLoss, customNLLLoss           # noqa: F401
from .host import HostPipeline                               # noqa: F401

__version__ = "0.1.0"
