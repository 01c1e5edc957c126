"""Acceptance of the opt-in VTM-side binary reader (tools/vtm_reader/pmp_partition_reader.h).

Build container only.  A copy of the vendored EncAppCfg.cpp is patched IN THE BUILD DIRECTORY (git-ignored; nothing from the
reference enters the repository): its two per-component text-parse loops (EncAppCfg.cpp:4301-4398) are replaced by two
pmp_read_partition() calls, the patched file is compiled with the flags of the stock build and linked against the stock
libraries into EncoderAppBin.  Each case is then encoded three times -- stock encoder + text files, patched encoder +
text files (fallback path), patched encoder + .bin files -- and the three bitstream MD5s must be equal.

    python tools/vtm_acceptance/run_binary_reader.py [cases_dir [case ...]]
"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools.vtm_acceptance import run as stock          # noqa: E402
from pmp_vvc_tip2023_b200 import partition_io           # noqa: E402

SRC = os.path.join(stock.VTM_SRC, "App", "EncoderApp")
BUILD = stock.BUILD


def build_patched():
    exe = os.path.join(BUILD, "bin", "EncoderAppBin")
    stock.build()
    if os.path.exists(exe):
        return exe
    pdir = os.path.join(BUILD, "patched")
    os.makedirs(pdir, exist_ok=True)
    text = open(os.path.join(SRC, "EncAppCfg.cpp"), encoding="latin-1").read()
    start = text.index("for (int frm = 0; frm < partitionFrameNum; frm++)", text.index("memory request finished"))
    end = text.rindex('cout << "parse finished" << endl;')      # the last one: an earlier one is commented out inside the loop
    call = ('if (pmp_read_partition(partitionMatPath + seqNameLuma, partitionFrameNum, partitionRow, partitionColumn, 0,\n'
            '        partitionHorMat, partitionVerMat, qtDepthMat, directionMat) < 0) exit(1);\n'
            '    if (pmp_read_partition(partitionMatPath + seqNameChroma, partitionFrameNum, partitionRow, partitionColumn, 1,\n'
            '        partitionHorMat, partitionVerMat, qtDepthMat, directionMat) < 0) exit(1);\n    ')
    # the stock code insists on opening the .txt files before parsing: the reader does its own opening (.bin, else .txt)
    head = text[:start].replace("if (!infileLuma)", "if (false)").replace("if (!infileChroma)", "if (false)")
    patched = '#include "%s"\n' % os.path.join(ROOT, "tools", "vtm_reader", "pmp_partition_reader.h") + head + call + text[end:]
    open(os.path.join(pdir, "EncAppCfg.cpp"), "w", encoding="latin-1").write(patched)
    cmds = subprocess.check_output(["ninja", "-t", "commands", "EncoderApp"], cwd=BUILD, text=True).splitlines()
    cc = [c for c in cmds if c.endswith("App/EncoderApp/EncAppCfg.cpp")][0]
    cc = cc.replace("-c " + os.path.join(SRC, "EncAppCfg.cpp"), "-I%s -c %s" % (SRC, os.path.join(pdir, "EncAppCfg.cpp")))
    cc = cc.replace("EncoderApp/CMakeFiles/EncoderApp.dir/EncAppCfg.cpp.o", "patched/EncAppCfg.o")
    subprocess.check_call(cc, shell=True, cwd=BUILD)
    link = [c for c in cmds if " -o bin/EncoderApp " in c][0].split("&& cd")[0]
    link = link.replace("EncoderApp/CMakeFiles/EncoderApp.dir/EncAppCfg.cpp.o", "patched/EncAppCfg.o").replace(
        "-o bin/EncoderApp ", "-o bin/EncoderAppBin ").replace("EncoderApp/CMakeFiles/EncoderApp.dir/link.d", "patched/link.d")
    subprocess.check_call(link.lstrip(": &"), shell=True, cwd=BUILD)
    return exe


def main():
    cases_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "vtm_cases")
    info = json.load(open(os.path.join(cases_dir, "cases.json")))
    names = sys.argv[2:] or ["pipe_192x128", "classD_416x240", "classC_832x480"]
    exe_stock, exe_bin = stock.build(), build_patched()
    from tools.vtm_acceptance.make_case import case_frames
    out = {}
    for name in names:
        ci = info[name]
        work = os.path.join(BUILD, "work_bin", name)
        os.makedirs(work, exist_ok=True)
        w, h, nf, y, u, v = case_frames(name)
        yuv = os.path.join(work, name + ".yuv")
        with open(yuv, "wb") as fp:
            for f in range(nf):
                fp.write(y[f].tobytes()); fp.write(u[f].tobytes()); fp.write(v[f].tobytes())
        tdir = os.path.join(cases_dir, name, "PartitionMat")
        bdir = os.path.join(work, "PartitionMat_bin")
        shutil.rmtree(bdir, ignore_errors=True)
        os.makedirs(bdir)
        for comp in ("Luma", "Chroma"):
            fn = "%s_%s_QP%d_PartitionMat" % (name, comp, ci["qp"])
            partition_io.write_partition_bin(os.path.join(bdir, fn + ".bin"), partition_io.text_to_values(os.path.join(tdir, fn + ".txt"), h, w), h, w)
        r_stock = stock.encode(exe_stock, yuv, tdir, name, ci, os.path.join(work, "stock_text"))
        r_fallback = stock.encode(exe_bin, yuv, tdir, name, ci, os.path.join(work, "patched_text"))
        r_bin = stock.encode(exe_bin, yuv, bdir, name, ci, os.path.join(work, "patched_bin"))
        out[name] = {"md5_stock_reader_text": r_stock["md5"], "md5_patched_reader_text": r_fallback["md5"],
                     "md5_patched_reader_bin": r_bin["md5"], "seconds": [r_stock["seconds"], r_fallback["seconds"], r_bin["seconds"]],
                     "all_equal": r_stock["md5"] is not None and r_stock["md5"] == r_fallback["md5"] == r_bin["md5"]}
    print(json.dumps(out, indent=1))
    return 0 if all(r["all_equal"] for r in out.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
