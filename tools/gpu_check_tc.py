#!/usr/bin/env python
"""Developer check (B200, via gpurun): tensor-core path parity vs oracle + timings."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import mediquery_rag_b200 as M

def parity(n, dtype, B, k, mask_frac=None, seed=0x5EED0003):
    x = O.synth_rows(seed, 0, n, 768)
    idx = M.Index(768, dtype); idx.add(x); idx.set_path(2)
    q, planted = O.planted_queries(0x0BADC0DE, B, seed, n, 768)
    st, _ = O.normalize_rows(x, dtype, 768); qn, _ = O.normalize_rows(q, "fp32", 768)
    mask = None
    if mask_frac is not None:
        mask = O.mask_from_bool(np.random.default_rng(1).random(n) < mask_frac)
    s, i = idx.search(q, k, mask)
    stt = idx.stats()
    es, ei = O.brute_force_topk(qn, st, k, mask)
    ok = np.array_equal(i, ei) and np.array_equal(s.view(np.uint32), es.view(np.uint32))
    print(f"tc parity n={n} {dtype} B={B} k={k} mask={mask_frac}: bit-exact={ok} path={stt['last_path']} fallbacks={stt['exact_fallbacks']}", flush=True)
    if not ok:
        bad = np.argwhere(i != ei)
        print("  mismatches:", len(bad), bad[:5].tolist())
        if len(bad): b = bad[0][0]; print("  got", i[b][:12], s[b][:6]); print("  exp", ei[b][:12], es[b][:6])
    idx.close()
    return ok

def timing(n, dtype, Bs, k=10, path=0):
    idx = M.Index(768, dtype, capacity_rows=n); idx.add_synthetic(0x5EED0002, 0, n); idx.set_path(path)
    out = []
    for B in Bs:
        q, planted = O.planted_queries(0x0BADC0DE, B, 0x5EED0002, n, 768)
        for _ in range(2): s, i = idx.search(q, k)
        okp = all(i[b, 0] == planted[b] for b in range(0, B, 2))
        ms = []
        for _ in range(5):
            idx.search(q, k); st = idx.stats(); ms.append((st["last_scan_ms"], st["last_search_ms"]))
        scan = float(np.median([m[0] for m in ms])); tot = float(np.median([m[1] for m in ms]))
        tf = 2.0 * n * 768 * B / (scan * 1e-3) / 1e12
        gb = n * 768 * (4 if dtype == "fp32" else 2) / 1e9 / (scan * 1e-3)
        out.append(dict(n=n, dtype=dtype, B=B, k=k, path=st["last_path"], scan_ms=round(scan, 3), search_ms=round(tot, 3), tflops=round(tf, 1), gbps=round(gb, 1), qps=round(B / tot * 1e3), planted_ok=okp, fallbacks=st["exact_fallbacks"]))
        print(out[-1], flush=True)
    idx.close()
    return out

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ok = True
    if which in ("all", "parity"):
        for args in [(8192, "bf16", 16, 10), (8192, "fp32", 16, 10), (50000, "bf16", 100, 10), (50000, "fp32", 130, 10), (100003, "bf16", 300, 100),
                     (100003, "fp32", 64, 32), (70000, "bf16", 64, 10, 0.3), (70000, "fp32", 33, 64, 0.05), (300000, "bf16", 512, 10)]:
            ok &= parity(*args)
        print("TC PARITY", "OK" if ok else "FAILED", flush=True)
    res = []
    if which in ("all", "time"):
        res += timing(1_000_000, "fp32", [16, 64, 128, 256])
        res += timing(1_000_000, "bf16", [16, 64, 256, 1024, 4096])
        res += timing(10_000_000, "bf16", [4096], k=100)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(dict(parity=bool(ok), timing=res), open("gpurun_out/gpu_check_tc.json", "w"), indent=1)
