"""Recipe for oracle/_ref/: the UNMODIFIED hot-path files of the reference, for timing the real reference on the GPU box.

    python oracle/install_ref.py          # build container only; needs /root/reference

`bench.py --impl reference` and the `cpu_baseline` leg time the reference's own `ReportGenerationModel.generate()` on the
host cores.  /root/reference does not exist on the GPU box, and the reference is pure Python without an installable
package layout (its setup.py's find_packages() finds nothing: there are no __init__.py files, the authors rely on
`pip install -e .`), so `pip install --target` yields an empty install.  This script therefore copies, byte for byte, the
eight files the path consists of into oracle/_ref/src/..., which is git-ignored (it never enters the
history) but NOT gpurun-ignored (it travels to the GPU box like a built .so).  Nothing is edited; oracle/ref_harness.py
imports the copies with the same 2 stubs + 2 patches it uses for /root/reference.  SHA-256 of every copied file is
written to oracle/_ref/MANIFEST.txt.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = [
    "src/full_model/report_generation_model.py",
    "src/object_detector/object_detector.py",
    "src/object_detector/custom_rpn.py",
    "src/object_detector/custom_roi_heads.py",
    "src/object_detector/image_list.py",
    "src/binary_classifier/binary_classifier_region_selection.py",
    "src/binary_classifier/binary_classifier_region_abnormal.py",
    "src/language_model/language_model.py",
]


def install(reference_root: str = "/root/reference") -> bool:
    if not os.path.isdir(os.path.join(reference_root, "src", "full_model")):
        return False
    lines = []
    for rel in FILES:
        src = os.path.join(reference_root, rel)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        lines.append("%s  %s" % (hashlib.sha256(open(dst, "rb").read()).hexdigest(), rel))
    for name in ("LICENSE", "NOTICE"):
        if os.path.exists(os.path.join(reference_root, name)):
            shutil.copyfile(os.path.join(reference_root, name), os.path.join(DST, name))
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as f:
        f.write("unmodified copies from ttanida/rgrg (see LICENSE / NOTICE); sha256  path\n" + "\n".join(lines) + "\n")
    return True


if __name__ == "__main__":
    ok = install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("oracle/_ref installed" if ok else "reference checkout not found; nothing installed")
