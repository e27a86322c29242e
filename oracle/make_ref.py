"""Recipe for oracle/_ref/: the REAL upstream hot-path primitives, so that they travel to the GPU box.

Upstream is pure Python with nothing to compile; the two modules the reference CPU leg needs (`metrics.py`:
edit_dist / evaluate, `CTCdecoder.py`: collapse_fn / CTCDecoder) are copied verbatim from /root/reference into
oracle/_ref/ (git-ignored, NOT gpurun-ignored -- exactly like a compiled oracle/_ref/*.so would be).  Nothing under
oracle/_ref/ is tracked, nothing in the product imports it; only oracle/refpath.py (test infrastructure) does.

    python oracle/make_ref.py            # no-op when /root/reference is absent (the GPU box uses the copied files)
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
FILES = ("metrics.py", "CTCdecoder.py")


def make():
    if not os.path.isdir(REF):
        return False
    os.makedirs(OUT, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(OUT, f))
    return True


if __name__ == "__main__":
    print("copied" if make() else "no /root/reference here; nothing to do")
