"""CPU side of the VTM acceptance harness: build the vendored VTM-10.0 EncoderApp (PMP fast-partition patches,
/root/reference, build container only) and encode the cases produced by make_case.py twice -- with OUR PartitionMat files
and with files the UNMODIFIED reference (Metrics.seq_post_process -> Map2Partition.get_sequence_partition_for_VTM) decodes
from the very same float maps (maps.npz) -- and compare files and bitstream MD5s.

    python tools/vtm_acceptance/run.py [cases_dir [case ...]]       -> JSON summary on stdout

Checks per case: EncoderApp parses the Luma/Chroma PartitionMat files (EncAppCfg.cpp:4234-4404), encodes all frames and
exits 0 with a non-empty bitstream; our files equal the reference-decoded files byte for byte; both encodes give the same
bitstream MD5.  For `pipe_192x128` the files are additionally compared with the committed golden files that the
reference's own nets + decode wrote (tests/golden).  Cases run in parallel worker processes (VTM is single-threaded).
"""
import contextlib
import hashlib
import io
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
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


def encode(exe, yuv, part_dir, name, info, work):
    os.makedirs(work, exist_ok=True)
    shutil.rmtree(os.path.join(work, "PartitionMat"), ignore_errors=True)
    shutil.copytree(part_dir, os.path.join(work, "PartitionMat"))
    cfg = os.path.join(work, name + ".cfg")
    with open(cfg, "w") as fp:
        fp.write("InputFile : %s\nInputBitDepth : 10\nFrameRate : 30\nFrameSkip : 0\nSourceWidth : %d\nSourceHeight : %d\n"
                 "FramesToBeEncoded : %d\nLevel : 6.2\n" % (yuv, info["width"], info["height"], info["frames"]))
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


def reference_decode(maps, info, name, out_dir):
    """PartitionMat files from the float maps through the unmodified reference post-process + decode."""
    import numpy as np
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self          # Metrics.py hard-codes .cuda(); build container has no GPU
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import Metrics as RefMetrics
    os.makedirs(out_dir, exist_ok=True)
    t0 = time.time()
    for comp in ("Luma", "Chroma"):
        path = os.path.join(out_dir, "%s_%s_QP%d_PartitionMat.txt" % (name, comp, info["qp"]))
        with contextlib.redirect_stdout(io.StringIO()):
            RefMetrics.seq_post_process(torch.from_numpy(np.array(maps[comp + "_qt"])), np.array(maps[comp + "_bt"]),
                                        np.array(maps[comp + "_dire"]), comp, info["frames"], info["width"], info["height"], path)
    return time.time() - t0


def run_case(cases_dir, name, info):
    import numpy as np
    from tools.vtm_acceptance.make_case import case_frames
    exe = build()
    cdir = os.path.join(cases_dir, name)
    work = os.path.join(BUILD, "work", name)
    os.makedirs(work, exist_ok=True)
    w, h, nf, y, u, v = case_frames(name)
    assert (w, h, nf) == (info["width"], info["height"], info["frames"])
    yuv = os.path.join(work, name + ".yuv")
    with open(yuv, "wb") as fp:
        for f in range(nf):
            fp.write(y[f].tobytes()); fp.write(u[f].tobytes()); fp.write(v[f].tobytes())
    maps = np.load(os.path.join(cdir, "maps.npz"))
    ref_dir = os.path.join(work, "ref_PartitionMat")
    t_ref = reference_decode(maps, info, name, ref_dir)
    same = True
    for comp in ("Luma", "Chroma"):
        fn = "%s_%s_QP%d_PartitionMat.txt" % (name, comp, info["qp"])
        same &= open(os.path.join(ref_dir, fn), "rb").read() == open(os.path.join(cdir, "PartitionMat", fn), "rb").read()
    res = encode(exe, yuv, os.path.join(cdir, "PartitionMat"), name, info, os.path.join(work, "ours"))
    ref = encode(exe, yuv, ref_dir, name, info, os.path.join(work, "ref"))
    res["accepted"] = res["rc"] == 0 and res["bitstream_bytes"] > 0 and res["pocs"] == info["frames"]
    res["reference_decode_seconds"] = round(t_ref, 1)
    res["files_identical_to_reference_decode_of_same_maps"] = bool(same)
    res["reference_files_md5"] = ref["md5"]
    res["md5_matches_reference_files"] = ref["md5"] == res["md5"] and ref["md5"] is not None
    res["blocks"] = int(maps["Luma_qt"].shape[0])
    res["decode_report"] = info.get("decode_report")
    if name == "pipe_192x128":
        gold_same = True
        for comp in ("Luma", "Chroma"):
            src = os.path.join(ROOT, "tests", "golden", "pipeline_%s_QP32_PartitionMat.txt" % comp)
            ours = os.path.join(cdir, "PartitionMat", "%s_%s_QP32_PartitionMat.txt" % (name, comp))
            gold_same &= open(src, "rb").read() == open(ours, "rb").read()
        res["files_identical_to_reference_golden"] = bool(gold_same)     # reference nets AND reference decode
    return res


def main():
    cases_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "vtm_cases")
    info = json.load(open(os.path.join(cases_dir, "cases.json")))
    if len(sys.argv) > 3 and sys.argv[2] == "--one":          # worker mode
        print(json.dumps(run_case(cases_dir, sys.argv[3], info[sys.argv[3]])))
        return 0
    names = sys.argv[2:] or list(info)
    build()
    procs = {n: subprocess.Popen([sys.executable, os.path.abspath(__file__), cases_dir, "--one", n], stdout=subprocess.PIPE,
                                 stderr=subprocess.PIPE, text=True) for n in names}
    out = {}
    for n, p in procs.items():
        so, se = p.communicate()
        try:
            out[n] = json.loads(so.strip().splitlines()[-1])
        except (ValueError, IndexError):
            out[n] = {"accepted": False, "error": se[-2000:]}
    print(json.dumps(out, indent=1))
    ok = all(r.get("accepted") and r.get("md5_matches_reference_files") for r in out.values())
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
