"""CPU side of the VTM acceptance harness: build the vendored VTM-10.0 EncoderApp (PMP fast-partition patches,
/root/reference, build container only) and encode the cases produced by make_case.py with OUR PartitionMat files.

    python tools/vtm_acceptance/run.py [cases_dir]       -> JSON summary on stdout

Checks per case: EncoderApp parses the Luma/Chroma PartitionMat files (EncAppCfg.cpp:4234-4404), encodes all frames and
exits 0 with a non-empty bitstream; for the `pipe_192x128` case the files are also compared byte for byte with the ones
the reference's own Python wrote (tests/golden) and both encodes must give the same bitstream MD5.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
VTM_SRC = "/root/reference/codec/vtm10.0-source-with-pmp-fast-alg"
BUILD = os.path.join(ROOT, "tools", "vtm_acceptance", "_build")
INTRA_CFG = "/root/reference/codec/demo/cfg/encoder_intra_vtm.cfg"


def build():
    exe = os.path.join(BUILD, "bin", "EncoderApp")
    if os.path.exists(exe):
        return exe
    os.makedirs(BUILD, exist_ok=True)
    subprocess.check_call(["cmake", "-G", "Ninja", os.path.join(ROOT, "tools", "vtm_acceptance"), "-DVTM_SRC=" + VTM_SRC], cwd=BUILD)
    subprocess.check_call(["ninja", "-j", str(max(1, (os.cpu_count() or 2) - 1))], cwd=BUILD)
    return exe


def encode(exe, case_dir, name, info, work):
    os.makedirs(work, exist_ok=True)
    shutil.rmtree(os.path.join(work, "PartitionMat"), ignore_errors=True)
    shutil.copytree(os.path.join(case_dir, "PartitionMat"), os.path.join(work, "PartitionMat"))
    cfg = os.path.join(work, name + ".cfg")
    with open(cfg, "w") as fp:
        fp.write("InputFile : %s\nInputBitDepth : 10\nFrameRate : 30\nFrameSkip : 0\nSourceWidth : %d\nSourceHeight : %d\n"
                 "FramesToBeEncoded : %d\nLevel : 4\n" % (os.path.join(case_dir, name + ".yuv"), info["width"], info["height"], info["frames"]))
    t0 = time.time()
    r = subprocess.run([exe, "-c", cfg, "-c", INTRA_CFG, "-f", str(info["frames"]), "-ts", "1", "-q", str(info["qp"]),
                        "-b", "enc.bin", "-o", "", "--SEIDecodedPictureHash=1"], cwd=work, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    dt = time.time() - t0
    bin_path = os.path.join(work, "enc.bin")
    size = os.path.getsize(bin_path) if os.path.exists(bin_path) else 0
    md5 = hashlib.md5(open(bin_path, "rb").read()).hexdigest() if size else None
    poc = [ln for ln in r.stdout.splitlines() if ln.startswith("POC")]
    return {"rc": r.returncode, "seconds": round(dt, 1), "bitstream_bytes": size, "md5": md5, "pocs": len(poc),
            "tail": r.stdout.splitlines()[-3:] if r.returncode else []}


def main():
    cases_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "vtm_cases")
    info = json.load(open(os.path.join(cases_dir, "cases.json")))
    exe = build()
    out = {}
    for name, ci in info.items():
        cdir = os.path.join(cases_dir, name)
        res = encode(exe, cdir, name, ci, os.path.join(BUILD, "work", name))
        res["accepted"] = res["rc"] == 0 and res["bitstream_bytes"] > 0 and res["pocs"] == ci["frames"]
        if name == "pipe_192x128":
            # same frames through the files the reference's own Python wrote
            gold = os.path.join(BUILD, "work", name + "_golden_case")
            os.makedirs(os.path.join(gold, "PartitionMat"), exist_ok=True)
            same = True
            for comp in ("Luma", "Chroma"):
                src = os.path.join(ROOT, "tests", "golden", "pipeline_%s_QP32_PartitionMat.txt" % comp)
                dst = os.path.join(gold, "PartitionMat", "%s_%s_QP32_PartitionMat.txt" % (name, comp))
                shutil.copy(src, dst)
                ours = os.path.join(cdir, "PartitionMat", "%s_%s_QP32_PartitionMat.txt" % (name, comp))
                same &= open(src, "rb").read() == open(ours, "rb").read()
            shutil.copy(os.path.join(cdir, name + ".yuv"), os.path.join(gold, name + ".yuv"))
            ref = encode(exe, gold, name, ci, os.path.join(BUILD, "work", name + "_golden"))
            res["files_identical_to_reference"] = same
            res["md5_matches_reference_files"] = ref["md5"] == res["md5"] and ref["md5"] is not None
        out[name] = res
    print(json.dumps(out, indent=1))
    return 0 if all(r["accepted"] for r in out.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
