"""Golden vectors that ARE reference outputs: runs the reference's own spoiling script
(/root/reference/scripts/generateDataset.py, unmodified, as a subprocess) on two small clean g2o files and commits
outputs next to the inputs (per case: the sha256 and size of the whole output file, and verbatim the outlier lines the
script appended — the head of the file is the input's vertex and edge records). tests/test_spoil_reference.py then
requires ipc_b200.spoil.spoil_g2o to write byte-identical files for the same input and options — this row (the outlier injector of the
Monte-Carlo protocol) is pinned against the reference itself, not against our own restatement.

The script is the one Python piece of the reference that runs in this container (the C++ path needs g2o, absent).
/root/reference does not exist on the GPU box: only this generator reads it, the tests read the committed files.
Run:  python tests/golden/make_golden_spoil.py
"""
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "spoil")
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ipc_b200 import g2o, synth  # noqa: E402

REF = "/root/reference/scripts/generateDataset.py"

# name -> (input, extra command-line options of the reference script, the same as keyword arguments of spoil_g2o)
CASES = {
    "2d_default":   ("clean_2d.g2o", ["-n", "40", "--seed", "7"], dict(outliers=40, seed=7)),
    "2d_local_grp": ("clean_2d.g2o", ["-n", "12", "--seed", "11", "-g", "3", "-l"], dict(outliers=12, seed=11, groupsize=3, local=True)),
    "2d_info1":     ("clean_2d.g2o", ["-n", "9", "--seed", "3", "--information", "42.7"], dict(outliers=9, seed=3, information="42.7")),
    "2d_info_full": ("clean_2d.g2o", ["-n", "5", "--seed", "5", "--information", "10,1,0,20,2,30"], dict(outliers=5, seed=5, information="10,1,0,20,2,30")),
    "2d_perfect":   ("clean_2d.g2o", ["-n", "6", "--seed", "9", "-p"], dict(outliers=6, seed=9, perfect_match=True)),
    "2d_none":      ("clean_2d.g2o", ["-n", "0", "--seed", "1"], dict(outliers=0, seed=1)),
    "3d_default":   ("clean_3d.g2o", ["-n", "30", "--seed", "13"], dict(outliers=30, seed=13)),
    "3d_info1":     ("clean_3d.g2o", ["-n", "4", "--seed", "2", "--information", "100"], dict(outliers=4, seed=2, information="100")),
    "3d_perfect":   ("clean_3d.g2o", ["-n", "3", "--seed", "4", "-p", "-g", "2"], dict(outliers=3, seed=4, perfect_match=True, groupsize=2)),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    g2o.write_g2o(synth.make_clean("intel", 0.1), os.path.join(OUT, "clean_2d.g2o"))
    g2o.write_g2o(synth.make_clean("sphere", 0.04), os.path.join(OUT, "clean_3d.g2o"))
    meta = {}
    for name, (src, opts, kw) in CASES.items():
        dst = os.path.join(OUT, f"ref_{name}.tmp")
        subprocess.check_call([sys.executable, REF, "-i", os.path.join(OUT, src), "-o", dst] + opts, stdout=subprocess.DEVNULL)
        data = open(dst, "rb").read()
        os.remove(dst)
        n_in = sum(1 for ln in open(os.path.join(OUT, src)) if not ln.startswith("#"))
        tail = data.decode().splitlines(keepends=True)[n_in:]
        with open(os.path.join(OUT, f"ref_{name}.tail"), "w") as f:
            f.writelines(tail)
        meta[name] = dict(input=src, reference_options=opts, spoil_kwargs=kw, sha256=hashlib.sha256(data).hexdigest(), n_bytes=len(data),
                          n_outlier_lines=len(tail))
        print(name, len(data), "bytes", len(tail), "outlier lines")
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
