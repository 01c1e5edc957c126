"""Random-shape stress of the TC conv engine through pmp_selftest_conv (GPU box): layer shapes, batches and epilogue
variants the nets do not use, each checked against the exact fp32 SIMT conv; every group of cases runs in a child
process with a timeout so that a hang is reported instead of stalling the run.

    python tools/fuzz_conv.py [n_cases [seed]]
"""
import ctypes
import json
import random
import subprocess
import sys

sys.path.insert(0, ".")


def cases(n, seed):
    rng = random.Random(seed)
    out = []
    while len(out) < n:
        cin = rng.choice([3, 8, 16, 32, 64, 128])
        cout = rng.choice([8, 16, 32, 64])
        k = rng.choice([1, 3, 3, 3, 5])
        hw = rng.choice([8, 16, 16, 32, 32, 64])
        if hw == 64 and cin * k * k > 64 * 25:
            continue
        big = rng.random() < 0.5
        b = rng.choice([1, 2, 3, 5, 7, 74, 75, 149]) if not big else rng.choice([600, 1201, 2400, 3000])
        if hw == 64:
            b = min(b, 600)
        fl = rng.choice([0, 1, 3, 5, 7, 1, 3])
        mode = rng.choice(["plain", "plain", "fused", "hpool"])
        if mode == "fused" and k > 1:
            fl = (fl & ~2) | (1 << 17) | (rng.choice([3, 16, 32, 64]) << 20)
        elif mode == "hpool" and k > 1:
            fl = (fl & ~4) | (1 << 18)
        out.append((cin, cout, k, hw, b, fl))
    return out


def child(group):
    from pmp_vvc_tip2023_b200 import _lib
    h = _lib.Handle.get(0)
    L = _lib.lib()
    for cfg in group:
        cin, cout, k, hw, b, fl = cfg
        me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
        rc = L.pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, fl, ctypes.byref(me), ctypes.byref(am), ctypes.byref(t1), ctypes.byref(t2))
        if rc == -3 or (rc and b"not supported" in (L.pmp_last_error() or b"")):
            print(json.dumps({"cfg": cfg, "status": "unsupported"}), flush=True)
            continue
        ok = rc == 0 and me.value <= 2e-5 * max(am.value, 1.0)
        print(json.dumps({"cfg": cfg, "status": "ok" if ok else "FAIL", "rc": rc, "err": me.value, "absmax": am.value,
                          "msg": (L.pmp_last_error() or b"").decode() if rc else ""}), flush=True)
        if rc == -2:
            return


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(json.loads(sys.argv[2]))
        return 0
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    cs = cases(n, seed)
    bad = 0
    done = 0
    unsupported = 0
    for i in range(0, len(cs), 8):
        group = cs[i:i + 8]
        try:
            r = subprocess.run([sys.executable, __file__, "--child", json.dumps(group)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                               text=True, timeout=120)
            lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
        except subprocess.TimeoutExpired as e:
            lines = [json.loads(l) for l in (e.stdout or b"").decode().splitlines() if l.startswith("{")]
            hung = group[len(lines)] if len(lines) < len(group) else None
            print("HANG", hung, flush=True)
            bad += 1
        for l in lines:
            done += 1
            unsupported += l["status"] == "unsupported"
            if l["status"] == "FAIL":
                bad += 1
                print("FAIL", l, flush=True)
    print("cases run %d of %d (%d not supported by the TC engine), failures/hangs %d" % (done, len(cs), unsupported, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
