"""Top-level stand-in for upstream CTCdecoder.py: put <repo> and <repo>/policy-gradient-asr_b200/dropin on
sys.path ahead of the upstream checkout and `import CTCdecoder` resolves here."""
from pgasr_b200.CTCdecoder import *            # noqa: F401,F403
from pgasr_b200 import CTCdecoder as _impl
globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith("__")})
