#!/usr/bin/env python3
"""Staged bring-up on a real B200: every stage runs in its own subprocess under a timeout so a
faulting kernel cannot take the remaining stages (or the box) with it.

    python tools/gpu_bringup.py            # all stages, log to gpurun_out/bringup.log
    python tools/gpu_bringup.py <stage> [args]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402


def _imports():
    import torch
    import normalisr_oracle as orc
    import nsr_testlib as tl
    from normalisr_b200 import association, engine, synth
    return torch, orc, tl, association, engine, synth


def stage_pvalue():
    torch, orc, tl, association, engine, synth = _imports()
    ctx = engine.context(0)
    d = np.load(os.path.join(ROOT, "tests/golden/pvalue_kat.npz"))
    a, r2, P = d["a"], d["r2"], d["P"]
    got = engine.pvalue(ctx, torch.from_numpy(r2).cuda(), a[:, 0]).cpu().numpy()
    m = P >= 1e-300
    rel = np.abs(got - P)[m] / P[m]
    print("pvalue KAT: max rel err %.3e, tail max %.3e" % (rel.max(), got[~m].max()))
    assert rel.max() < 1e-9


def _resid_case(torch, tl, association, engine, rows, n, nc, n_slices, had, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(rows, n)) * rng.uniform(0.2, 5, size=(rows, 1)) + rng.normal(size=(rows, 1)) * 3
    x[rng.random((rows, n)) < 0.7] *= 0.01          # heavy tails
    x[0, n // 3] = 1e5                              # one giant outlier: exercises the re-quantisation pass
    if rows > 3:
        x[2] = 3.25                                 # constant row: residual is rounding noise when nc > 0
    dc = np.concatenate([rng.normal(size=(max(nc - 1, 0), n)), np.ones((1 if nc else 0, n))])
    Qt, rank, W = association.covariate_basis(dc)
    ctx = engine.context(0)
    engine.set_option("hadamard", int(had))
    S = engine.residualize(ctx, torch.from_numpy(x).cuda(), torch.from_numpy(Qt).cuda() if rank else None,
                           n_slices, keep_coef=True)
    torch.cuda.synchronize()
    z = tl.residual(x, Qt)
    zt = tl.hadamard128(z) if had else np.pad(z, ((0, 0), (0, S.n_pad - n)))
    var = (z ** 2).mean(1)
    got_var = S.var.cpu().numpy()
    got = engine.unslice(ctx, S).cpu().numpy()
    q = S.quantum.cpu().numpy()
    ok_rows = var > 1e-20 * (x ** 2).mean(1)       # rows that are pure rounding noise are not comparable
    err = (np.abs(got - zt) / q[:, None])[ok_rows]
    var, got_var = var[ok_rows], got_var[ok_rows]
    sl = S.slices.cpu().numpy()
    info = dict(rows=rows, n=n, nc=nc, S=n_slices, had=had,
                var_rel=float(np.abs(got_var - var).max() / var.max()),
                quant_err_in_quanta=float(err.max()),
                absmax_digit=int(np.abs(sl.astype(int)).max()),
                top_digit_max=int(np.abs(sl[0].astype(int)).max()),
                coef_err=float(np.abs(S.coef.cpu().numpy() - x @ Qt.T).max()) if rank else 0.0,
                gram_err=float(np.abs(got[ok_rows] @ got[ok_rows].T - z[ok_rows] @ z[ok_rows].T).max() /
                               np.abs(z[ok_rows] @ z[ok_rows].T).max()))
    print(json.dumps(info))
    assert info["var_rel"] < 1e-12 and info["quant_err_in_quanta"] <= 0.5 + 1e-6 and info["top_digit_max"] <= 127
    return info


def stage_residual():
    torch, orc, tl, association, engine, synth = _imports()
    for (rows, n, nc) in [(37, 300, 4), (200, 1000, 9), (64, 1280, 0), (9, 4097, 13), (130, 128, 1)]:
        for S in (3, 4):
            for had in (1, 0):
                _resid_case(torch, tl, association, engine, rows, n, nc, S, had)
    engine.set_option("hadamard", 1)


def _golden_coex(eng_name, precision):
    torch, orc, tl, association, engine, synth = _imports()
    from normalisr_b200 import normalisr as norm
    eng = {"simt": engine.ENGINE_SIMT, "umma": engine.ENGINE_UMMA}[eng_name]
    worst = {}
    for case in ["coex_chain", "coex_tail", "coex_rankdef", "coex_nocov"]:
        g = np.load(os.path.join(ROOT, "tests/golden", case + ".npz"))
        ka = {"dimreduce": int(g["dimreduce"])} if "dimreduce" in g.files else {}
        P, dot, var = norm.coex(g["dt"], g["dc"], engine=eng, precision=precision, **ka)
        r = dot / np.sqrt(np.outer(var, var))
        rr = g["dot"] / np.sqrt(np.outer(g["var"], g["var"]))
        m = g["P"] >= 1e-300
        relp = np.abs(P - g["P"])[m] / g["P"][m]
        worst[case] = dict(dr=float(np.abs(r - rr).max()), relP=float(relp.max()),
                           var=float(np.abs(var - g["var"]).max()), diagP=float(np.abs(np.diag(P)).max()),
                           tail=float(P[~m].max()) if (~m).any() else 0.0)
        print(case, json.dumps(worst[case]))
    return worst


def stage_simt_golden():
    for prec in ("default", "fast", "precise"):
        print("precision", prec)
        _golden_coex("simt", prec)


def stage_umma_golden(precision="default", kblock="128"):
    torch, orc, tl, association, engine, synth = _imports()
    engine.set_option("umma_kblock", int(kblock))
    print("umma golden precision", precision, "kblock", kblock)
    _golden_coex("umma", precision)


def stage_umma_vs_simt(precision="default", kblock="128", rows="700", n="5000", pair="1"):
    """bit-identical outputs from the two engines on identical digit planes"""
    torch, orc, tl, association, engine, synth = _imports()
    rows, n = int(rows), int(n)
    engine.set_option("umma_kblock", int(kblock))
    engine.set_option("umma_pair", int(pair))
    if os.environ.get("NSR_EPI_WARPS"):
        engine.set_option("epi_warps", int(os.environ["NSR_EPI_WARPS"]))
    ctx = engine.context(0)
    p = synth.host_problem(7, rows, n)
    Qt, rank, W = association.covariate_basis(p["dc"])
    S, prods = engine.PRESETS[precision]
    A = engine.residualize(ctx, torch.from_numpy(p["dt"]).cuda(), torch.from_numpy(Qt).cuda(), S)
    outs = []
    for eng in (engine.ENGINE_SIMT, engine.ENGINE_UMMA):
        P = torch.full((rows, rows), -1.0, dtype=torch.float64, device="cuda")
        D = torch.full((rows, rows), -1.0, dtype=torch.float64, device="cuda")
        engine.contract(ctx, engine.MODE_COEX, A, A, engine.coex_tiles(rows), (n - 1 - rank) / 2, P, D, prods, eng)
        torch.cuda.synchronize()
        outs.append((P.cpu().numpy(), D.cpu().numpy()))
    same_p = np.array_equal(outs[0][0], outs[1][0])
    same_d = np.array_equal(outs[0][1], outs[1][1])
    print("pair %s precision %s kblock %s rows %d n %d: P identical %s, dot identical %s, max|dP| %.3e max|dD| %.3e" % (
        pair, precision, kblock, rows, n, same_p, same_d, np.abs(outs[0][0] - outs[1][0]).max(),
        np.abs(outs[0][1] - outs[1][1]).max()))
    # exact integer check of the SIMT engine against numpy int64
    sl = A.slices.cpu().numpy()[:, :64]
    tot = tl.kept_products_sum(sl, sl, {6: 4, 8: 5, 10: 5}[prods])
    q = A.quantum.cpu().numpy()[:64]
    want = tot.astype(np.float64) * np.outer(q, q) / n
    np.fill_diagonal(want, 0.0)
    print("SIMT vs numpy int64: max rel %.3e" % (np.abs(outs[0][1][:64, :64] - want).max() / np.abs(want).max()))
    assert same_p and same_d


def _time_ms(torch, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


def stage_perf(rows="5000", n="10000", precision="default", kblock="128", check="1", pair="1"):
    torch, orc, tl, association, engine, synth = _imports()
    rows, n = int(rows), int(n)
    engine.set_option("umma_kblock", int(kblock))
    engine.set_option("umma_pair", int(pair))
    ctx = engine.context(0)
    t0 = time.time()
    p = synth.device_problem(1002, rows, n, "cuda")
    torch.cuda.synchronize()
    print("generated %dx%d on device in %.1fs" % (rows, n, time.time() - t0))
    Qt, rank, W = association.covariate_basis(p["dc"].cpu().numpy())
    Qd = torch.from_numpy(Qt).cuda()
    S, prods = engine.PRESETS[precision]
    out = engine.Sliced(rows, n, S, "cuda")
    tr = _time_ms(torch, lambda: engine.residualize(ctx, p["dt"], Qd, S, out=out))
    P = torch.empty((rows, rows), dtype=torch.float64, device="cuda")
    D = torch.empty((rows, rows), dtype=torch.float64, device="cuda")
    tiles = engine.coex_tiles(rows)
    dof = (n - 1 - rank) / 2
    tc = _time_ms(torch, lambda: engine.contract(ctx, engine.MODE_COEX, out, out, tiles, dof, P, D, prods, engine.ENGINE_UMMA))
    pairs = rows * (rows - 1) / 2
    ops = 2.0 * prods * len(tiles) * 128 * 128 * out.n_pad
    print(json.dumps(dict(rows=rows, n=n, precision=precision, kblock=int(kblock), pair=int(pair), residualize_ms=tr, contract_ms=tc,
                          pairs_per_s=pairs / ((tr[0] + tc[0]) * 1e-3), int8_tops=ops / (tc[0] * 1e-3) / 1e12,
                          alg_tflops=2.0 * n * pairs / (tc[0] * 1e-3) / 1e12,
                          resid_GBs=16.0 * rows * n / (tr[0] * 1e-3) / 1e9)))
    if int(check):
        P2 = torch.empty_like(P)
        D2 = torch.empty_like(D)
        sub = tiles[:: max(1, len(tiles) // 40)]
        P.fill_(-1); D.fill_(-1); P2.fill_(-1); D2.fill_(-1)
        engine.contract(ctx, engine.MODE_COEX, out, out, sub, dof, P, D, prods, engine.ENGINE_UMMA)
        engine.contract(ctx, engine.MODE_COEX, out, out, sub, dof, P2, D2, prods, engine.ENGINE_SIMT)
        torch.cuda.synchronize()
        print("subset check vs SIMT: identical P %s dot %s" % (bool((P == P2).all()), bool((D == D2).all())))


STAGES = [
    ("pvalue", [], 300),
    ("residual", [], 300),
    ("simt_golden", [], 300),
    ("umma_golden", ["default", "128"], 180),
    ("umma_golden", ["default", "64"], 180),
    ("umma_golden", ["fast", "128"], 180),
    ("umma_golden", ["precise", "128"], 180),
    ("umma_vs_simt", ["default", "128"], 240),
    ("umma_vs_simt", ["default", "64"], 240),
    ("umma_vs_simt", ["fast", "128"], 240),
    ("umma_vs_simt", ["precise", "128"], 240),
    ("perf", ["5000", "10000", "default", "128"], 300),
    ("perf", ["5000", "10000", "default", "64"], 300),
    ("perf", ["5000", "10000", "fast", "128"], 300),
    ("perf", ["5000", "10000", "precise", "128"], 300),
    ("perf", ["8192", "65536", "default", "128", "0"], 400),
    ("perf", ["8192", "65536", "default", "64", "0"], 400),
]


QUICK = [
    ("residual", [], 300),
    ("umma_vs_simt", ["default", "128", "700", "5000", "1"], 120),
    ("umma_vs_simt", ["default", "128", "300", "1000", "1"], 120),
    ("umma_vs_simt", ["fast", "128", "700", "5000", "1"], 120),
    ("umma_vs_simt", ["precise", "128", "700", "5000", "1"], 120),
    ("umma_vs_simt", ["default", "128", "1333", "3001", "1"], 120),
    ("umma_golden", ["default", "128"], 180),
    ("umma_vs_simt", ["default", "128", "700", "5000", "0"], 120),
    ("umma_vs_simt", ["precise", "64", "700", "5000", "0"], 120),
    ("perf", ["5000", "10000", "default", "128", "1", "1"], 300),
    ("perf", ["5000", "10000", "default", "128", "1", "0"], 300),
    ("perf", ["8192", "65536", "default", "128", "0", "1"], 400),
    ("perf", ["8192", "65536", "default", "128", "0", "0"], 400),
    ("perf", ["8192", "65536", "fast", "128", "0", "1"], 400),
    ("perf", ["8192", "65536", "precise", "128", "0", "1"], 400),
]


def main():
    global STAGES
    if len(sys.argv) > 1 and sys.argv[1] == "--quick":
        STAGES = QUICK
    elif len(sys.argv) > 1:
        globals()["stage_" + sys.argv[1]](*sys.argv[2:])
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "bringup.log"), "a")
    summary = []
    for name, args, tmo in STAGES:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name] + args, capture_output=True,
                               text=True, timeout=tmo)
            rc, out = r.returncode, r.stdout + r.stderr
        except subprocess.TimeoutExpired as e:
            rc, out = -999, "TIMEOUT\n" + str(e.stdout or "") + str(e.stderr or "")
        line = "=== %s %s rc=%d (%.1fs)" % (name, " ".join(args), rc, time.time() - t0)
        print(line)
        print(out[-3000:])
        log.write(line + "\n" + out + "\n")
        log.flush()
        summary.append((name, args, rc))
    print("SUMMARY")
    for s in summary:
        print(s)


if __name__ == "__main__":
    main()
