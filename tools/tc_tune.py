#!/usr/bin/env python
"""Sweep the launch-plan variants of the tcgen05 point-wise GEMM for every 1x1 convolution of the network
(batch 32 @ 640x640 by default) and report the fastest correct one per (K, N).

    python tools/tc_tune.py --out gpurun_out/tc_tune.jsonl [--batch 32] [--only K,N]

Every variant is checked against an fp64 matmul on sampled rows before it is timed (cf_debug_pw_gemm_time).  Results are
appended to the JSONL file as they are produced; a variant that kills the CUDA context (bounded mbarrier wait -> trap) is
recorded as started-but-not-finished and skipped when the tool is re-run, so a wrapper loop can resume the sweep.
The winners go into tc_tuned_table (csrc/k_pw_tc.cuh)."""
import argparse
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("lightweight-face-detection-centernet_b200")

LIN, SWISH, RES = 0, 1, 2


def layers(batch, h=640, w=640):
    """(name, M, K, N, epi) of every point-wise convolution, model/centernet.py:211-239."""
    blocks = [(32, 16, 1, 3, 1), (16, 24, 6, 3, 2), (24, 24, 6, 3, 1), (24, 32, 6, 5, 2), (32, 32, 6, 5, 1), (32, 64, 6, 3, 2),
              (64, 64, 6, 3, 1), (64, 96, 6, 5, 1), (96, 96, 6, 5, 1), (96, 160, 6, 5, 2), (160, 160, 6, 5, 1), (160, 320, 6, 3, 1)]
    out = []
    hh, ww = h // 2, w // 2
    for i, (cin, cout, t, k, s) in enumerate(blocks):
        hid = cin * t
        if t != 1:
            out.append((f"b{i}.exp", batch * hh * ww, cin, hid, SWISH))
        hh, ww = hh // s, ww // s
        out.append((f"b{i}.proj", batch * hh * ww, hid, cout, RES if (cin == cout and s == 1) else LIN))
    out.append(("clast", batch * hh * ww, 320, 24, SWISH))
    for j, c in enumerate((96, 32, 24)):
        hh, ww = hh * 2, ww * 2
        out.append((f"up{j + 1}", batch * hh * ww, c, 24, LIN))
    return out


def variants(K, N):
    n32 = (N + 31) // 32 * 32
    v = []
    for nc in range(32, min(128, n32) + 1, 32):
        for direct in (0, 1, 2):
            for atmem in ((1, 0) if nc <= 96 else (0,)):
                for rchunk in (1, 0):
                    v.append({"CF_TC_NC": nc, "CF_TC_DIRECT": direct, "CF_TC_ATMEM": atmem, "CF_PWN": 0, "CF_TC_RCHUNK": rchunk})
        for atmem in ((1, 0) if nc <= 96 else (0,)):
            v.append({"CF_TC_NC": nc, "CF_TC_DIRECT": 0, "CF_TC_ATMEM": atmem, "CF_PWN": 0, "CF_TC_STG": 4})
        if nc <= 64:
            v.append({"CF_TC_NC": nc, "CF_TC_DIRECT": 0, "CF_TC_ATMEM": 1, "CF_PWN": 0, "CF_TC_NACC": 3})
        if nc <= 64 and K <= 32 and n32 <= nc:
            v.append({"CF_TC_NC": nc, "CF_PWN": 1})
    return v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tc_tune.jsonl"))
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--only", default="")
    ap.add_argument("--default-only", action="store_true", help="time only the library's default plan of each layer")
    a = ap.parse_args()
    lib = pkg._lib.load()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    done = set()
    if os.path.exists(a.out):
        for line in open(a.out):
            try:
                r = json.loads(line)
                done.add(r["key"])
            except Exception:
                pass
    f = open(a.out, "a")
    seen = set()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for name, M, K, N, epi in layers(a.batch):
        if (M, K, N, epi) in seen:
            continue
        seen.add((M, K, N, epi))
        if a.only and a.only != f"{K},{N}":
            continue
        g = torch.Generator(device="cuda").manual_seed(K * 1000 + N)
        A = torch.randn(M, K, device="cuda", generator=g)
        Wm = np.ascontiguousarray((np.random.RandomState(K + N).randn(K, N) / np.sqrt(K)).astype(np.float32))
        R = torch.randn(M, N, device="cuda", generator=g) if epi == RES else None
        out = torch.empty(M, N, device="cuda")
        idx = torch.cat([torch.randint(M, (4096,), device="cuda", generator=g), torch.arange(M - 300, M, device="cuda"),
                         torch.arange(0, 300, device="cuda")])
        ref = A[idx].double() @ torch.from_numpy(Wm).cuda().double()
        if epi == SWISH:
            ref = ref * torch.sigmoid(ref)
        if epi == RES:
            ref = ref + R[idx].double()
        scale = ref.abs().max().item()
        iters = 20 if M <= 204800 else 8
        for v in [{}] + ([] if a.default_only else variants(K, N)):  # {} = the library's current default plan
            key = f"{name}|{M}|{K}|{N}|{epi}|" + ",".join(f"{k}={v[k]}" for k in sorted(v))
            if key in done:
                continue
            f.write(json.dumps({"key": key, "status": "started"}) + "\n")
            f.flush()
            os.fsync(f.fileno())
            for k in ("CF_TC_NC", "CF_TC_DIRECT", "CF_TC_ATMEM", "CF_PWN", "CF_TC_RCHUNK", "CF_TC_STG", "CF_TC_NACC"):
                os.environ.pop(k, None)
            for k, val in v.items():
                os.environ[k] = str(val)
            out.fill_(float("nan"))
            ms = C.c_float()
            desc = C.create_string_buffer(256)
            rc = lib.cf_debug_pw_gemm_time(1, epi, C.c_void_p(A.data_ptr()), C.c_void_p(Wm.ctypes.data), C.c_void_p(out.data_ptr()),
                                           M, K, N, C.c_void_p(R.data_ptr()) if R is not None else None, st, iters, C.byref(ms), desc, 256)
            rec = {"key": key, "name": name, "M": M, "K": K, "N": N, "epi": epi, "variant": v, "rc": rc, "plan": desc.value.decode()}
            if rc != 0:
                rec["status"] = "error"
                rec["error"] = lib.cf_last_error().decode(errors="replace")
                print("ERR", key, rec["error"], flush=True)
                f.write(json.dumps(rec) + "\n")
                f.flush()
                if rc == -2:  # CUDA error: the context is gone, let the wrapper restart us
                    sys.exit(3)
                continue
            err = ((out[idx].double() - ref).abs().max().item()) / scale
            full = bool(torch.isfinite(out).all().item())
            rec.update(status="ok" if (err < 2e-5 and full) else "wrong", ms=ms.value, rel_err=err, finite=full)
            print(f"{name:9s} M={M:8d} K={K:4d} N={N:4d} {rec['status']:5s} {ms.value * 1e3:8.1f} us  err {err:.1e}  {rec['plan']}  {v}", flush=True)
            f.write(json.dumps(rec) + "\n")
            f.flush()
        del A, out, R
    f.close()
    # summary: best correct variant per layer
    best = {}
    for line in open(a.out):
        r = json.loads(line)
        if r.get("status") != "ok":
            continue
        k = (r["name"], r["M"], r["K"], r["N"])
        if not r["variant"]:
            best.setdefault(k, {})["default"] = r
        if k not in best or "best" not in best[k] or r["ms"] < best[k]["best"]["ms"]:
            best.setdefault(k, {})["best"] = r
    print("\n| layer | M | K | N | default us | best us | best plan |\n|---|---|---|---|---|---|---|")
    tot_d = tot_b = 0.0
    for (name, M, K, N), d in best.items():
        dflt = d.get("default", d["best"])
        tot_d += dflt["ms"]
        tot_b += d["best"]["ms"]
        print(f"| {name} | {M} | {K} | {N} | {dflt['ms'] * 1e3:.1f} | {d['best']['ms'] * 1e3:.1f} | {d['best']['plan']} {d['best']['variant']} |")
    print(f"| sum (distinct shapes) | | | | {tot_d * 1e3:.0f} | {tot_b * 1e3:.0f} | |")


if __name__ == "__main__":
    main()
