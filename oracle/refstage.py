"""Stage an UNMODIFIED copy of the reference under baseline/_ref/ so that it travels to the GPU box.

Oracle / test infrastructure only (see oracle/__init__.py).  /root/reference exists only in the build
container; baseline/_ref/ is git-ignored (never part of the history) but NOT gpurun-ignored, so the GPU
box gets the reference's own kzg.py / fft_ff.py / plonk / marlin / main.py and fixtures byte for byte.
The reference is plain Python files, not an installable package, so a file copy is the whole "install"
(`pip install --target baseline/_ref /root/reference` has nothing to build: there is no setup.py / pyproject).
`__graft_entry__.build()` calls stage(); `bench.py --impl reference`, oracle/refrun.py and the GPU tests call
root() to find whichever copy exists.
"""
import filecmp
import os
import shutil

SOURCE = "/root/reference"
STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
_KEEP = (".py", ".pkl", ".md")


def _files(root):
    for d, dirs, files in os.walk(root):
        dirs[:] = [x for x in dirs if not x.startswith(".") and x != "__pycache__"]
        for f in files:
            if f.endswith(_KEEP) or f == "LICENSE":
                yield os.path.relpath(os.path.join(d, f), root)


def stage():
    """Copy /root/reference -> baseline/_ref (only when the source is mounted).  Returns the staged root or None."""
    if not os.path.isfile(os.path.join(SOURCE, "kzg.py")):
        return STAGED if os.path.isfile(os.path.join(STAGED, "kzg.py")) else None
    for rel in _files(SOURCE):
        src, dst = os.path.join(SOURCE, rel), os.path.join(STAGED, rel)
        if os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return STAGED


def root():
    """Directory holding the reference's own sources: the mounted tree if present, else the staged copy, else None."""
    for r in (SOURCE, STAGED):
        if os.path.isfile(os.path.join(r, "kzg.py")):
            return r
    return None


def verify_staged():
    """True iff every staged file is byte-identical to the mounted reference (build container only)."""
    if not os.path.isfile(os.path.join(SOURCE, "kzg.py")):
        return None
    return all(os.path.isfile(os.path.join(STAGED, rel)) and filecmp.cmp(os.path.join(SOURCE, rel), os.path.join(STAGED, rel), shallow=False)
               for rel in _files(SOURCE))
