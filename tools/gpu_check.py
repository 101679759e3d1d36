#!/usr/bin/env python
"""Developer check run on a B200 through gpurun: parity vs the oracle on small shapes + quick scan timings."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import mediquery_rag_b200 as M

def parity(n, dim, dtype, B, k, mask_frac=None, seed=0x5EED0002):
    x = O.synth_rows(seed, 0, n, dim)
    idx = M.Index(dim, dtype)
    idx.add(x)
    q, planted = O.planted_queries(0x0BADC0DE, B, seed, n, dim)
    st, _ = O.normalize_rows(x, dtype, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    mask = None
    if mask_frac is not None:
        rng = np.random.default_rng(1)
        mask = O.mask_from_bool(rng.random(n) < mask_frac)
    s, i = idx.search(q, k, mask)
    es, ei = O.brute_force_topk(qn, st, k, mask)
    stored = idx.get_rows(np.arange(min(n, 64)))
    rows_ok = np.array_equal(stored, O.stored_as_f32(st)[:min(n, 64), :dim])
    ok = np.array_equal(i, ei) and np.array_equal(s.view(np.uint32), es.view(np.uint32))
    print(f"parity n={n} dim={dim} {dtype} B={B} k={k} mask={mask_frac}: ids+scores bit-exact={ok} rows_bit_exact={rows_ok}", flush=True)
    if not ok:
        bad = np.argwhere(i != ei)
        print("  first mismatches:", bad[:5].tolist(), i[bad[0][0]][:k], ei[bad[0][0]][:k], s[bad[0][0]][:4], es[bad[0][0]][:4])
    idx.close()
    return ok and rows_ok

def timing(n, dtype, Bs, k=10):
    idx = M.Index(768, dtype, capacity_rows=n)
    t = time.time(); idx.add_synthetic(0x5EED0002, 0, n); t_add = time.time() - t
    out = []
    for B in Bs:
        q, _ = O.planted_queries(0x0BADC0DE, B, 0x5EED0002, n, 768)
        for _ in range(3): idx.search(q, k)
        ms = []
        for _ in range(10):
            idx.search(q, k); st = idx.stats(); ms.append((st["last_scan_ms"], st["last_search_ms"], st["last_scan_passes"]))
        scan = np.median([m[0] for m in ms]); tot = np.median([m[1] for m in ms]); passes = ms[0][2]
        gb = n * 768 * (4 if dtype == "fp32" else 2) * passes / 1e9
        out.append(dict(n=n, dtype=dtype, B=B, passes=passes, scan_ms=round(scan, 4), search_ms=round(tot, 4), gbps=round(gb / scan * 1e3, 1), qps=round(B / tot * 1e3, 1)))
        print(out[-1], flush=True)
    print("add_synthetic s:", round(t_add, 3))
    idx.close()
    return out

if __name__ == "__main__":
    ok = True
    for args in [(154, 768, "fp32", 1, 4), (154, 768, "fp32", 3, 5), (1000, 768, "bf16", 5, 10), (20000, 768, "fp32", 9, 10),
                 (20000, 768, "bf16", 8, 10, 0.3), (5000, 384, "fp32", 2, 100), (3000, 100, "bf16", 3, 100, 0.01),
                 (40000, 768, "fp32", 2, 1000), (7, 768, "fp32", 2, 10), (100000, 768, "bf16", 4, 10, 0.0001)]:
        ok &= parity(*args)
    print("PARITY", "OK" if ok else "FAILED", flush=True)
    res = timing(1_000_000, "fp32", [1, 2, 4, 8]) + timing(1_000_000, "bf16", [1, 2, 4, 8])
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(dict(parity=bool(ok), timing=res), open("gpurun_out/gpu_check.json", "w"), indent=1)
