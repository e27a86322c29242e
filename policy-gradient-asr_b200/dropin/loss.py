"""Top-level stand-in for upstream loss.py: put <repo> and <repo>/policy-gradient-asr_b200/dropin on
sys.path ahead of the upstream checkout and `import loss` resolves here."""
from pgasr_b200.loss import *            # noqa: F401,F403
from pgasr_b200 import loss as _impl
globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith("__")})
