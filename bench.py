#!/usr/bin/env python
"""Headline benchmark: images/sec at 640x640 batch inference (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (u8 batch -> network -> sigmoid/clamp -> path-C decode -> [B,100,6]
boxes, + for N>1 the NCCL all-gather of the box lists) over one synthetic batch of 32 images per GPU
(configs[1]; at N=8 this is configs[2], batch-256 sharded 32/GPU => weak scaling).

`value`  : inputs resident in HBM, timed with CUDA events on the launching stream, max over ranks.
`e2e`    : the same step through the C-ABI host entry points cf_submit_topk_host / cf_wait_host with
           pinned HOST buffers (every batch's H2D and D2H inside the timed region, double buffered).
`roofline`: the dominant kernel class, its own algorithmic bytes (cf_work_model x batch: a fused MBConv launch reads its
           block input and writes its block output) / CUDA-event time of that class's launches, against
           MEASURED_PEAKS.json; the special-function-unit roofline of the fused launches beside it.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference from baseline/_ref on the host cores (oracle port as fall-back).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG_NAME = "lightweight-face-detection-centernet_b200"
WEIGHTS = os.path.join(ROOT, "tests", "golden", "weights_e100.npz")
METRIC = "images/sec at 640x640 batch inference"
UNIT = "images/s"
H = W = 640
PER_GPU_BATCH = 32
K_TOP = 100
CLASS_NAMES = {1: "pointwise_gemm", 2: "depthwise", 3: "stem", 4: "heads", 5: "decode_topk", 6: "fused_blocks"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="images per GPU per step")
    ap.add_argument("--pw", type=int, default=int(os.environ.get("CENTERFACE_B200_PW", -1)),
                    help="engine: 0 fp32 SIMT, 1 tcgen05 3xTF32 + fused MBConv blocks, 2 tcgen05 1xTF32, 3 tcgen05 3xTF32 layer-wise "
                         "(default: library default)")
    ap.add_argument("--cpu-sample", type=int, default=PER_GPU_BATCH, help="images per step of the cpu_baseline leg of the B200 arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(batch, world):
    """config.workload, the same string in both arms (the driver compares the arms' configs)."""
    return (f"batch-{batch} 640x640 per GPU ({'configs[1]' if world == 1 else 'configs[2] sharding'}), "
            f"u8 BGR input, network + sigmoid/clamp + path-C top-{K_TOP} decode" + (", NCCL all-gather of [B,100,6] boxes" if world > 1 else ""))


def synthetic_batch(n, seed):
    """SURVEY.md 8d throughput inputs: seeded u8 uniform noise [n,H,W,3] (backbone cost is content
    independent; random-noise heat-maps still exercise the full peak/top-k decode)."""
    rng = np.random.RandomState(seed)
    return rng.randint(0, 256, size=(n, H, W, 3), dtype=np.uint8)


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path on all host cores.  baseline/_ref holds the UNMODIFIED
# reference files (staged by oracle/make_ref.py in the build container, git-ignored, shipped to the GPU box); when it is
# absent the oracle port (oracle/centerface_oracle.py, the same ATen calls) stands in and the line says kind = "port".
# ---------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def load_reference():
    """(net, CenterFace class on CPU, ctdet_decode, mean, std) of the unmodified reference, or None."""
    if not os.path.exists(os.path.join(REF_DIR, "centerface.py")):
        return None
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))  # stub mlconfig / torchsummary
    sys.path.insert(0, REF_DIR)
    cwd = os.getcwd()
    os.chdir(REF_DIR)  # the checkpoint path is cwd-relative (centerface.py:23)
    try:
        import model.centernet as mc
        mc.ghost_net = mc.efficientnet_b0  # lets centerface_ext import; its ctdet_decode is model-independent
        import centerface as ref_cf
        import centerface_ext as ref_ext

        class CPUCenterFace(ref_cf.CenterFace):  # centerface.py:16-27 hard-codes .cuda(): restated for the CPU
            def __init__(self, height, width, landmarks=True):
                self.landmarks = landmarks
                self.net = mc.efficientnet_b0()
                self.cuda = False
                self.net.load_state_dict(torch.load("weight/model_epoch_100.pt", map_location="cpu", weights_only=True))
                self.net.eval()
                self.img_h_new, self.img_w_new, self.scale_h, self.scale_w = self.transform(height, width)

        cf = CPUCenterFace(H, W)
    finally:
        os.chdir(cwd)
    return cf.net, cf, ref_ext.ctdet_decode, ref_cf.CenterFace.mean, ref_cf.CenterFace.std


def cpu_reference_run(steps, warmup, sample, with_call=False):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    u8 = synthetic_batch(sample, 0)
    ref = load_reference()
    extra = {}
    if ref is not None:
        net, cf, ctdet_decode, mean, std = ref
        kind = "reference"

        def step():
            with torch.no_grad():
                x = np.stack([((im / 255. - mean) / std).astype(np.float32).transpose(2, 0, 1) for im in u8])  # centerface.py:32-35
                out = net(torch.from_numpy(x))[0]                                                             # :41
                hm = out["hm"].sigmoid_().clamp(1e-4, 1 - 1e-4)                                               # :43
                return ctdet_decode(hm, out["wh"], out["reg"], K=K_TOP)                                       # centerface_ext.py:52-82
        if with_call:  # the drop-in call itself (resize, normalise, forward, path A, Python NMS), F5 set at 640x640
            import contextlib
            import io
            import cv2
            imgs = [cv2.resize(cv2.imread(os.path.join(REF_DIR, "imgs", f"{n}.jpg")), (W, H)) for n in ("1", "17", "2", "27", "8")]
            ts = []
            with contextlib.redirect_stdout(io.StringIO()):
                for k in range(13):
                    t0 = time.perf_counter()
                    cf(imgs[k % 5], threshold=0.35)
                    ts.append(time.perf_counter() - t0)
            extra["call_ms_per_image_f5_640"] = round(float(np.median(ts[3:])) * 1e3, 1)
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import centerface_oracle as O
        sd = O.load_weights(WEIGHTS)
        kind = "port"

        def step():
            x = torch.from_numpy(np.stack([O.normalize_u8(i) for i in u8]))  # centerface.py:32-34
            o = O.forward(sd, x)                                             # model/centernet.py:263-280
            dets, _ = O.ctdet_decode(O.sigmoid_clamp(o["hm"]), o["wh"], o["reg"], K=K_TOP)  # centerface_ext.py:52-82
            return dets

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    cb = {"value": sample * steps / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
          "sample": f"{steps} steps x {sample} images of the 640x640 batch-{PER_GPU_BATCH} workload "
                    f"(normalise + forward + sigmoid/clamp + ctdet_decode K={K_TOP}), "
                    f"{'unmodified reference from baseline/_ref' if kind == 'reference' else 'oracle port'}, torch {torch.__version__} CPU, "
                    f"{cores} logical cores"}
    cb.update(extra)
    return cb, dt / steps * 1e3


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, a.steps), max(0, a.warmup)
    cb, ms = cpu_reference_run(steps, warm, PER_GPU_BATCH)  # the whole batch-32 step, as many steps as asked for
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(PER_GPU_BATCH, 1), "global_batch": PER_GPU_BATCH, "h": H, "w": W, "parallelism": "dp1",
                       "where": "host CPU, all cores (the reference's own CPU path; it has no other)"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md)
# ---------------------------------------------------------------------------------------------
class Clocks:
    """Samples SM clock, power and throttle reasons every 100 ms from ONE streaming nvidia-smi process while the
    timed regions run (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                c = [x.strip() for x in line.strip().split(",")]
                if len(c) >= 7:
                    self.rows.append(c)
        except Exception:
            pass

    def __enter__(self):
        self.t.start()
        time.sleep(0.3)
        return self

    def __exit__(self, *exc):
        time.sleep(0.15)
        if self.proc is not None:
            self.proc.terminate()
        self.t.join(timeout=5)

    def summary(self):
        rows = []
        for r in self.rows:
            try:
                rows.append((float(r[0]), float(r[1]), float(r[2]), r[3:7]))
            except ValueError:
                continue
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(r[0] for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3][i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": rows[0][1], "reasons": reasons,
                "power_w_max": max(r[2] for r in rows), "samples": len(rows)}


BLOCKS = [(32, 16, 1, 3, 1), (16, 24, 6, 3, 2), (24, 24, 6, 3, 1), (24, 32, 6, 5, 2), (32, 32, 6, 5, 1), (32, 64, 6, 3, 2),
          (64, 64, 6, 3, 1), (64, 96, 6, 5, 1), (96, 96, 6, 5, 1), (96, 160, 6, 5, 2), (160, 160, 6, 5, 1), (160, 320, 6, 3, 1)]


TOPK_SMEM_MAX_HW = 51200  # csrc/k_decode.cuh


def launch_table(h, w, fused=(), dwp=(), own=False):
    """(name, algorithmic bytes per image) of every launch of one forward + path-C decode, in launch order -- the same
    layer-wise accounting as cf_work_model (un-padded input once + output once + residual / low-res re-reads, fp32).
    A block in `fused` is ONE launch credited with the layer-wise bytes of the three launches it replaces; a block in `dwp` keeps its
    expand launch (if it has one) and runs depth-wise + projection as one launch credited with the bytes of those two.
    own=True: a fused launch counts only what it has to move itself (its input once + its output once + the residual)."""
    out = []
    hh, ww = h // 2, w // 2
    out.append(("stem 3->32 s2", h * w * 3 + hh * ww * 32 * 4))
    for i, (cin, cout, t, k, s) in enumerate(BLOCKS):
        hid = cin * t
        ho, wo = hh // s, ww // s
        res = cout if (cin == cout and s == 1) else 0
        rows = []
        if t != 1:
            rows.append((f"b{i} expand {cin}->{hid}", hh * ww * (cin + hid) * 4))
        rows.append((f"b{i} dw{k}x{k} s{s} {hid}ch", (hh * ww + ho * wo) * hid * 4))
        rows.append((f"b{i} project {hid}->{cout}" + (" +res" if res else ""), ho * wo * (hid + cout + res) * 4))
        if i in fused:
            by = (hh * ww * cin + ho * wo * (cout + res)) * 4 if own else sum(b for _, b in rows)
            rows = [(f"b{i} fused MBConv {cin}->{hid}->{cout} k{k} s{s}" + (" +res" if res else ""), by)]
        elif i in dwp:
            by = (hh * ww * hid + ho * wo * (cout + res)) * 4 if own else sum(b for _, b in rows[-2:])
            rows = rows[:-2] + [(f"b{i} fused dw{k}x{k} s{s} + project {hid}->{cout}" + (" +res" if res else ""), by)]
        out += rows
        hh, ww = ho, wo
    out.append(("conv_last 320->24", hh * ww * (320 + 24) * 4))
    for j, c in enumerate((96, 32, 24)):
        lo = hh * ww
        hh, ww = hh * 2, ww * 2
        out.append((f"up{j + 1} {c}->24 (IDAUp)", (hh * ww * (c + 24) + lo * 24) * 4))
    out.append(("heads 3x3 24->15", hh * ww * (24 + 16) * 4))
    if hh * ww <= TOPK_SMEM_MAX_HW:  # one launch: the peak keep runs on the top-k kernel's shared copy of the map
        out.append(("peak mask + top-k + gather", hh * ww * 5 * 4 + 100 * 6 * 4))
    else:
        out.append(("peak mask", hh * ww * 2 * 4))
        out.append(("top-k + gather", hh * ww * 4 + 100 * 6 * 4))
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG_NAME)
    L = pkg._lib
    sh = importlib.import_module(PKG_NAME + ".sharding")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {a.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this arm has no CPU fallback (use --impl reference for the CPU path)")
    # host side of the end-to-end path: this rank's pinned buffers on the GPU's own NUMA node (restored before the CPU legs)
    aff_old, aff_new = (None, None) if os.environ.get("CF_BENCH_BIND", "1") == "0" else sh.bind_host_to_gpu(local)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if rank == 0:
        pkg.build()
    if world > 1:
        dist.barrier()
    B = a.batch
    pw = a.pw if a.pw >= 0 else L.CF_PW_TCGEN05
    eng = pkg.Engine(WEIGHTS, max_batch=B, max_h=H, max_w=W, device=local, pw_engine=pw)
    dev = torch.device(f"cuda:{local}")

    # rotating input buffers (4 x 39 MB > 126 MB L2); each step also streams ~12 GB of activations
    n_rot = 4
    host = [torch.from_numpy(synthetic_batch(B, 1000 * rank + i)).pin_memory() for i in range(n_rot)]
    devin = [h.to(dev) for h in host]
    out_dets = torch.empty((B, K_TOP, 6), dtype=torch.float32).pin_memory()
    out_inds = torch.empty((B, K_TOP), dtype=torch.int32).pin_memory()
    n_total = B * world

    def step_resident(i):
        eng.forward(devin[i % n_rot])
        dets, _ = eng.decode_topk(K_TOP)
        return sh.gather_detections(dets, n_total=n_total)  # NCCL all-gather of the final box list (N>1)

    # end-to-end outputs: with N > 1 every rank receives the gathered [N*B,100,6] list (the exchange step runs inside the library:
    # ncclAllGather on the compute stream right behind the decode kernel, then ONE device-to-host copy)
    outs = [(torch.empty((n_total, K_TOP, 6), dtype=torch.float32).pin_memory(),
             torch.empty((B, K_TOP), dtype=torch.int32).pin_memory()) for _ in range(2)]
    if world > 1:
        eng.comm_init()

    def e2e_loop(steps):
        """`steps` batches through the C-ABI host entry points, double buffered: the H2D of batch i+1
        overlaps the kernels of batch i; every batch's inputs are copied from pinned host memory and its
        boxes (N > 1: every rank's boxes) are read back to the host inside the timed region."""
        for i in range(steps):
            if world > 1:
                eng.submit_topk_gather_host(host[i % n_rot], K_TOP, outs[i % 2][0], outs[i % 2][1])
            else:
                eng.submit_topk_host(host[i % n_rot], K_TOP, outs[i % 2][0], outs[i % 2][1])
            if i >= 1:
                eng.wait_host()  # results of submission i - 1 are now in outs[(i - 1) % 2]
        eng.wait_host()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        sync()
        l0 = eng.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        sync()
        dev_ms = e0.elapsed_time(e1)
        t = torch.tensor([dev_ms, wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), eng.launches - l0

    warm = max(a.warmup, 3)
    with Clocks(local) as clk:
        dev_ms, _, launches = timed(step_resident, a.steps, warm)
        e2e_steps = max(4, min(a.steps, 20))
        e2e_loop(3)
        sync()
        t0 = time.perf_counter()
        e2e_loop(e2e_steps)
        torch.cuda.synchronize()
        e2e_wall = time.perf_counter() - t0
        sync()
        t = torch.tensor([e2e_wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_wall_ms = t[0].item()
    clocks = clk.summary()

    ms_per_step = dev_ms / a.steps
    value = n_total / (ms_per_step * 1e-3)
    e2e_value = n_total / (e2e_wall_ms / e2e_steps * 1e-3)

    # per-class roofline, measured live with CUDA events on the launching stream (rank 0's GPU)
    hbm, tf, peak_src = peaks()
    classes = {}
    eng.forward(devin[0])
    eng.decode_topk(K_TOP)
    torch.cuda.synchronize()
    for cls, name in CLASS_NAMES.items():
        ms, n = eng.time_class(cls, iters=5)
        by, fl = L.work_model(H, W, L.CF_IN_U8_HWC, cls, pw)
        if n == 0:
            continue
        # Algorithmic bytes of a kernel = SURVEY.md 8(d): the kernel's OWN un-padded input read once + output written once
        # (+ residual re-read).  For a fused MBConv launch that is the block's input and output -- the hidden tensors never
        # reach HBM -- so its HBM fraction is small by construction and the bound that explains it is the special-function
        # pipe (`fused_blocks_sfu` below).  `layerwise_bytes` = the bytes of the launches it replaces, reported beside it
        # (round 1's schedule moved them; `layerwise_equiv_GBps` can exceed the HBM peak: that is the saving, not bandwidth).
        by_lw = by
        if cls == L.CLS_FUSED:
            lw = lambda c: L.work_model(H, W, L.CF_IN_U8_HWC, c, L.CF_PW_TCGEN05_LAYERWISE)[0] - L.work_model(H, W, L.CF_IN_U8_HWC, c, pw)[0]  # noqa: E731
            by_lw = lw(L.CLS_PW) + lw(L.CLS_DW)
        classes[name] = {"ms_per_step": ms, "launches": n, "alg_bytes": by * B, "layerwise_bytes": by_lw * B, "alg_flops": fl * B,
                         "gbs": by * B / (ms * 1e-3) / 1e9, "gbs_lw": by_lw * B / (ms * 1e-3) / 1e9, "tflops": fl * B / (ms * 1e-3) / 1e12}
    net_ms = sum(c["ms_per_step"] for k, c in classes.items())
    for c in classes.values():
        c["share"] = c["ms_per_step"] / net_ms
    top = max(classes, key=lambda k: classes[k]["ms_per_step"])
    tc = classes[top]
    roofline = {"kernel": top, "bound": "hbm", "achieved": tc["gbs"], "peak": hbm, "unit": "GB/s", "frac": tc["gbs"] / hbm,
                "traffic": None, "peak_source": peak_src, "launches": tc["launches"], "ms": tc["ms_per_step"],
                "alg_bytes_per_step": tc["alg_bytes"], "layerwise_bytes_per_step": tc["layerwise_bytes"],
                "layerwise_equiv_GBps": tc["gbs_lw"], "share_of_step": tc["share"],
                "tensor_frac_of_sustained_bf16": tc["tflops"] / tf,
                "classes": {k: {"ms": round(v["ms_per_step"], 4), "share": round(v["share"], 4), "GBps": round(v["gbs"], 1),
                                "frac_hbm": round(v["gbs"] / hbm, 4), "layerwise_equiv_GBps": round(v["gbs_lw"], 1),
                                "TFLOPs": round(v["tflops"], 2), "launches": v["launches"]}
                            for k, v in classes.items()}}
    # The fused MBConv kernels move a fraction of the layer-wise bytes: what bounds them is the SM's special-function pipe -- the
    # reference's Swish = one exponential + one reciprocal per hidden element, counted as 2 MUFU ops at 16 lanes/clk/SM
    # (layer1.0's kernel shares one reciprocal among four values, i.e. issues 1.25: the algorithmic count stays 2).  Every
    # hidden element once at input resolution (expand Swish, not for layer0 where t = 1) and once at output resolution.
    try:
        fz = classes.get("fused_blocks")
        if fz:
            hid = {0: (32, 2, 1, False), 1: (96, 2, 2, True), 2: (144, 4, 1, True)}  # block: hidden channels, input stride, dw stride, has expand
            mask_f, mask_d = L.fused_blocks(pw), L.dwp_blocks(pw)
            sw = 0
            for i, (c, div, st, ex) in hid.items():
                pin, pout = (H // div) * (W // div), (H // (div * st)) * (W // (div * st))
                if i in mask_f:
                    sw += c * ((pin if ex else 0) + pout)
                elif i in mask_d:
                    sw += c * pout
            prop = torch.cuda.get_device_properties(0)
            mufu_peak = 16.0 * prop.multi_processor_count * (clocks.get("sm_mhz") or 1965.0) * 1e6  # ops/s
            ach = 2.0 * sw * B / (fz["ms_per_step"] * 1e-3)
            roofline["fused_blocks_sfu"] = {"bound": "sfu (MUFU ex2 + rcp per Swish)", "achieved": round(ach / 1e12, 3), "peak": round(mufu_peak / 1e12, 3),
                                            "unit": "Tops/s", "frac": round(ach / mufu_peak, 3), "swish_per_image": sw}
    except Exception as ex:  # instrumentation only
        roofline["fused_blocks_sfu_error"] = str(ex)
    try:  # measured DRAM bytes of the dominant class (ncu, committed under profiles/), if it matches this run
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r2.json")))
        if (tj.get("engine"), tj.get("batch"), tj.get("h"), tj.get("w")) == (pw, B, H, W):
            roofline["traffic"] = tj["dram_bytes_per_step"].get(top)
            roofline["traffic_source"] = "profiles/traffic_r2.json"
    except Exception:
        pass
    try:  # the individual launches, timed one by one (events between launches: no overlap of neighbouring kernels)
        ms_l, _ = eng.time_steps(5)
        tab = launch_table(H, W, L.fused_blocks(pw), L.dwp_blocks(pw), own=True)
        tab_lw = launch_table(H, W, L.fused_blocks(pw), L.dwp_blocks(pw))
        if len(tab) == len(ms_l):
            rows = [{"launch": n, "us": round(t * 1e3, 1), "alg_GB": round(by_l * B / 1e9, 4),
                     "GBps": round(by_l * B / (t * 1e-3) / 1e9, 1), "frac_hbm": round(by_l * B / (t * 1e-3) / 1e9 / hbm, 3)}
                    for (n, by_l), t in zip(tab, ms_l)]
            for r, (_, by_l), (_, by_w) in zip(rows, tab, tab_lw):
                if by_w != by_l:  # a fused launch: the bytes of the launches it replaces, for comparison with round 1's schedule
                    r["layerwise_GB"] = round(by_w * B / 1e9, 4)
            rows.sort(key=lambda r: -r["us"])
            roofline["top_launches"] = rows[:8]
            roofline["dominant_launch"] = rows[0]
    except Exception as ex:  # instrumentation only
        roofline["top_launches_error"] = str(ex)
    by_min, _ = L.work_model(H, W, L.CF_IN_U8_HWC, 0, pw)
    by_all, fl_all = L.work_model(H, W, L.CF_IN_U8_HWC, 0, L.CF_PW_TCGEN05_LAYERWISE)  # layer-wise algorithmic bytes
    # whole step: the bytes this schedule has to move (fused blocks: block input + output) against the HBM peak; the layer-wise
    # figure of SURVEY.md 8(d) (384.6 MB per image, what round 1 moved) beside it
    roofline["whole_step"] = {"alg_bytes_per_image": by_min, "layerwise_bytes_per_image": by_all, "alg_flops_per_image": fl_all,
                              "GBps": by_min * B / (ms_per_step * 1e-3) / 1e9,
                              "frac_hbm": by_min * B / (ms_per_step * 1e-3) / 1e9 / hbm,
                              "layerwise_equiv_GBps": by_all * B / (ms_per_step * 1e-3) / 1e9,
                              "TFLOPs": fl_all * B / (ms_per_step * 1e-3) / 1e12}

    gather_check = None
    if world > 1:
        # content of the exchange step: slice r of the gathered list must be rank r's own boxes -- on every rank
        own, _ = eng.detect_topk_host(host[0].numpy(), K_TOP)
        eng.submit_topk_gather_host(host[0], K_TOP, outs[0][0], outs[0][1])
        eng.wait_host()
        got = outs[0][0].numpy()
        ok_own = bool(np.array_equal(got[rank * B:(rank + 1) * B], own))
        digest = torch.tensor([float(np.abs(got[r * B:(r + 1) * B]).sum()) for r in range(world)], dtype=torch.float64, device=dev)
        ref = digest.clone()
        dist.broadcast(ref, src=0)
        flags = torch.tensor([1.0 if ok_own else 0.0, 1.0 if bool(torch.equal(ref, digest)) else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        gather_check = {"own_slice_equal_on_every_rank": bool(flags[0].item() == 1.0),
                        "gathered_list_equal_on_every_rank": bool(flags[1].item() == 1.0), "ranks": world}
        if not (gather_check["own_slice_equal_on_every_rank"] and gather_check["gathered_list_equal_on_every_rank"]):
            raise SystemExit(f"bench.py: the all-gathered box list is wrong: {gather_check}")

    if aff_old is not None and aff_new != aff_old:
        os.sched_setaffinity(0, aff_old)  # the CPU legs use every host core
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_baseline, _ = cpu_reference_run(3, 1, a.cpu_sample, with_call=True)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": warm,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {0: "f32", 1: "tf32x3", 2: "tf32", 3: "tf32x3", 6: "tf32x3/tf32"}[pw], "data": "synthetic",
                "config": {"workload": workload_name(B, world), "global_batch": n_total, "h": H, "w": W, "pw_engine": pw,
                           "l2": f"{n_rot} rotating input batches ({n_rot * B * H * W * 3 / 1e6:.0f} MB) and "
                                 f"{by_min * B / 1e9:.1f} GB of activation traffic per step, both > 126 MB L2",
                           "parallelism": f"dp{world}"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * H * W * 3,
                        "d2h_bytes_per_step": n_total * K_TOP * 6 * 4 + B * K_TOP * 4,
                        "api": ("cf_submit_topk_gather_host (ncclAllGather inside the library, one D2H of the gathered list)" if world > 1
                                else "cf_submit_topk_host") + " / cf_wait_host (pinned host buffers, double buffered)",
                        "steps": e2e_steps,
                        "host_cpus_bound_to_gpu_numa_node": (len(aff_new) if aff_new is not None else None)},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
        if gather_check is not None:
            line["gather_check"] = gather_check
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
