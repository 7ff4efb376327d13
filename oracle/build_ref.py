"""Recipe for oracle/_ref/: packs the files of the UNMODIFIED reference that its CPU path executes into ONE archive.

TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py).  The reference is pure Python and cannot be installed as a package; the
GPU box has no /root/reference.  ``__graft_entry__.build()`` therefore runs this recipe in the build container: the reference's
own modules for the hot path travel to the GPU box inside ``oracle/_ref/reference_hotpath.tar.gz`` (git-ignored like every
build output, shipped with the gpurun snapshot like the built ``.so``), and ``bench.py --impl reference`` / the
``cpu_baseline`` leg unpack it into a temporary directory and run the reference's OWN code there
(models/yolo.py ``Model`` on torch-cpu, utils/general.py ``non_max_suppression``, tracker/bytetrack.py ``ByteTrack.update``
through oracle/refshim.py, which only injects the missing third-party names).  Nothing is edited; no reference source enters
the repository history.

    python oracle/build_ref.py            # no-op (keeps an existing archive) when /root/reference is absent
"""
import io
import os
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT_DIR, "reference_hotpath.tar.gz")
REF = os.environ.get("B2T_REFERENCE_ROOT", "/root/reference")
# the hot path's modules and what they import at module level
FILES = ["tracker/basetrack.py", "tracker/bytetrack.py", "tracker/botsort.py", "tracker/kalman_filter.py", "tracker/matching.py",
         "tracker/reid_models/__init__.py", "tracker/reid_models/deepsort_reid.py",
         "cfg/deploy/yolov7-w6.yaml", "cfg/deploy/yolov7-tiny.yaml"]
DIRS = ["models", "utils"]          # top-level *.py only


def build(verbose=True):
    if not os.path.isdir(os.path.join(REF, "tracker")):
        if verbose:
            print("oracle/_ref: %s not present, keeping %s" % (REF, "the existing archive" if os.path.exists(ARCHIVE) else "nothing"))
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    names = list(FILES)
    for d in DIRS:
        names += sorted(os.path.join(d, f) for f in os.listdir(os.path.join(REF, d)) if f.endswith(".py"))
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz") as tar:
        for n in names:
            tar.add(os.path.join(REF, n), arcname=n)
    with open(ARCHIVE, "wb") as f:
        f.write(buf.getvalue())
    if verbose:
        print("oracle/_ref: packed %d reference files (%d KB) into %s" % (len(names), len(buf.getvalue()) // 1024, ARCHIVE))
    return ARCHIVE


def unpack(dst):
    """Extracts the archive into dst (a temporary directory chosen by the caller); returns dst or None when there is no archive."""
    if not os.path.exists(ARCHIVE):
        return None
    with tarfile.open(ARCHIVE, "r:gz") as tar:
        tar.extractall(dst)
    return dst


if __name__ == "__main__":
    build()
