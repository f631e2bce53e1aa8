"""Stage the UNMODIFIED reference for the CPU arm of bench.py (`--impl reference`, `cpu_baseline.kind = "reference"`).

The reference (HotzingTone/DX-VAE) is three Python scripts with no packaging, importing dgl and mido, neither of which is
installable offline, so `pip install --target baseline/_ref /root/reference` has nothing to install.  What the CPU arm
needs instead is the reference's own model.py / dxdata.py next to its dataset, byte for byte; they are copied into
baseline/_ref/ (git-ignored: reference sources never enter this repo's history; NOT gpurun-ignored, so the directory
travels to the GPU box) and run there under the ~150-line dgl/mido stand-in of oracle/shim.

Run in the build container (where /root/reference is mounted):  python baseline/stage_ref.py
__graft_entry__.build() calls it when the reference tree is present."""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("DXVAE_REFERENCE_ROOT", "/root/reference")
FILES = ["model.py", "dxdata.py", "README.md", os.path.join("DX_data", "DXDataset.bin"), os.path.join("DX_data", "Dexed_01.syx")]


def stage(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "model.py")):
        if verbose:
            print("stage_ref: no reference tree at %s (nothing staged)" % SRC)
        return False
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
    if verbose:
        print("stage_ref: staged %d files under %s" % (len(FILES), DST))
    return True


def staged():
    return all(os.path.isfile(os.path.join(DST, rel)) for rel in FILES[:2])


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
