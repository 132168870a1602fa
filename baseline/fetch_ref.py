"""Recipe for `baseline/_ref/`: the UNMODIFIED pure-Python modules of the reference's inference path, copied from /root/reference
where that exists (this container).  The reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref` cannot
be used; `baseline/_ref/` is git-ignored (no reference source enters the history) but travels to a GPU box with the snapshot,
where `bench.py --impl reference` would import it.  OPTIONAL and manual (`python baseline/fetch_ref.py`): nothing in the build, the
tests or the bench runs it -- by default no reference source is copied anywhere, and the reference arm uses /root/reference where
that exists and the bit-exact oracle port elsewhere."""
import os
import shutil

ROOT = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
FILES = ("KGnet.py", "postprocessing.py", "nms.py", "config.py")


def fetch():
    dst = os.path.join(ROOT, "_ref")
    if not all(os.path.exists(os.path.join(SRC, f)) for f in FILES):
        return dst if os.path.isdir(dst) else None
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(dst, f))
    return dst


if __name__ == "__main__":
    print(fetch())
