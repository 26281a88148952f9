#!/usr/bin/env python
"""bench.py — BASELINE.json metric on the BASELINE config, one JSON line on stdout (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...        (one rank per GPU; weak scaling: bs 32 per GPU, no data-path collective)

A step = one pass of the hot path (Model.forward + Detect decode, reference models/yolo.py) over one synthetic batch:
configs[1] "YOLOv3 640x640 bs=32 inference on 1 B200, synthetic input, random-init weights".
  value      images/s, inputs resident in HBM (fp32 NCHW), CUDA-graph replay, CUDA-event timing, max over ranks
  e2e        images/s through yolov3_b200.Pipeline with HOST uint8 images: H2D + forward + decode + NMS + D2H per step
  roofline   conv kernels (tensor bound): algorithmic conv FLOPs / event-timed conv_tc launch time, vs MEASURED_PEAKS
  cpu_baseline / --impl reference: the CPU oracle port of the reference forward (oracle/yolo_oracle.py, torch CPU ops,
             all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CFG = "yolov3.yaml"
IMG, BS = 640, 32
GFLOP_PER_IMG = 155.891          # 2*MAC over the 75 nn.Conv2d of yolov3.yaml @640
GFLOP_LAYER0 = 0.708             # layer 0 runs on CUDA cores (c_in=3); excluded from the tensor roofline
METRIC = "images/sec @640 bs32 YOLOv3"
UNIT = "images/s"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    # nvidia-smi needs ~0.2 s to deliver its first line and the timed region of the default run is shorter than that,
    # so the sampler is started before the warm-up, every line is stamped with its arrival time, and stop(t0, t1) keeps
    # the lines that arrived while the GPU ran this workload: the timed region [t0, t1] plus, when that holds fewer than
    # three, the identical untimed replays the caller appends right after it (same graph, same inputs, same clocks)

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append((time.time(), l)) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, l in self.lines:
            if t0 is not None and not (t0 + 0.05 <= ts <= t1 + 0.05):  # a line reports the ~100 ms before it arrived
                continue
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def aggregate(ms_local: float, dev) -> float:
    """Max over ranks of a locally event-timed duration (one process per GPU; NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms_local
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def build_model(device):
    import torch

    from yolov3_b200.model import Model

    torch.manual_seed(0)
    m = Model(CFG, device=device)
    # non-trivial BN statistics so that the fold is exercised
    g = torch.Generator().manual_seed(0)
    for k in list(m.params):
        if k.endswith("bn.weight"):
            m.params[k] = torch.rand(m.params[k].shape, generator=g) + 0.5
        elif k.endswith("bn.bias") or k.endswith("running_mean"):
            m.params[k] = torch.randn(m.params[k].shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            m.params[k] = torch.rand(m.params[k].shape, generator=g) + 0.5
    return m


_BEST_THREADS = None


def best_threads(om):
    """The reference leaves torch's intra-op thread count at its default; on many-core hosts that default can be far
    from the fastest setting for this conv stack, so the baseline uses the fastest of a few candidates (stated in the
    output)."""
    global _BEST_THREADS
    import torch

    if _BEST_THREADS is None:
        cores = os.cpu_count() or 1
        x = torch.rand(1, 3, IMG, IMG)
        best = (None, 1e30)
        with torch.inference_mode():
            for t in sorted({cores, max(1, cores // 2), max(1, cores // 4), min(cores, 16)}, reverse=True):
                torch.set_num_threads(t)
                om(x)
                t0 = time.perf_counter()
                om(x)
                dt = time.perf_counter() - t0
                if dt < best[1]:
                    best = (t, dt)
        _BEST_THREADS = best[0]
    torch.set_num_threads(_BEST_THREADS)
    return _BEST_THREADS


def cpu_forward_rate(n_img, iters, warmup=1):
    """Oracle port of the reference forward on the host cores; returns (images/s, threads used)."""
    import torch

    sys.path.insert(0, str(ROOT / "oracle"))
    import yolo_oracle as O

    om = O.OracleModel(ROOT / "yolov3_b200" / "cfg" / CFG, seed=0, fused=True)
    cores = best_threads(om)
    x = torch.rand(n_img, 3, IMG, IMG, generator=torch.Generator().manual_seed(1))
    with torch.inference_mode():
        for _ in range(warmup):
            om(x)
        t0 = time.perf_counter()
        for _ in range(iters):
            om(x)
        dt = time.perf_counter() - t0
    return n_img * iters / dt, cores


def run_reference(args, rank):
    """--impl reference: the reference's torch-CPU forward (oracle port) on rank 0's host cores."""
    if rank != 0:
        return
    n_img = 2
    rate, cores = cpu_forward_rate(n_img, 1, warmup=1)  # calibration
    budget = 150.0
    per_step = max(1, min(BS, int(budget / max(1, args.steps + args.warmup) * rate)))
    import torch

    sys.path.insert(0, str(ROOT / "oracle"))
    import yolo_oracle as O

    om = O.OracleModel(ROOT / "yolov3_b200" / "cfg" / CFG, seed=0, fused=True)
    cores = best_threads(om)
    x = torch.rand(per_step, 3, IMG, IMG, generator=torch.Generator().manual_seed(1))
    with torch.inference_mode():
        for _ in range(args.warmup):
            om(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            om(x)
        dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    # the reference's NMS on the host cores (config 5, conf 0.25 / IoU 0.45): oracle candidate pipeline (numpy) + the
    # reference's own torchvision.ops.nms call, 8 images
    nms_ref = None
    try:
        pred = O.synth_predictions(8, n_rows=25200, nc=80, seed=3)
        O.non_max_suppression(pred[:1], 0.25, 0.45, use_torchvision=True)
        t1 = time.perf_counter()
        O.non_max_suppression(pred, 0.25, 0.45, use_torchvision=True)
        t_nms = time.perf_counter() - t1
        nms_ref = {"conf0.25_iou0.45_single": {"input_boxes_per_s": 8 * 25200 / t_nms, "ms_per_batch_of_8": t_nms * 1e3},
                   "impl": "oracle candidate pipeline (numpy) + torchvision.ops.nms, the reference's call at general.py:733"}
    except Exception as e:  # torchvision missing: report why instead of failing the arm
        nms_ref = {"unavailable": repr(e)[:200]}
    sample = (f"{per_step} of the {BS} images of each step (fp32, fused BN, torch CPU ops, {cores} threads of "
              f"{os.cpu_count()} host cores: fastest of 4 thread counts tried)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"yolov3.yaml forward+decode {IMG}x{IMG}, CPU sample", "imgsz": IMG, "batch_per_step": per_step},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "nms": nms_ref,
        "gpu_launches": 0,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-op", default=None, help="write the per-launch timing table (JSON) to this path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from yolov3_b200 import _lib
    from yolov3_b200.pipeline import Pipeline

    model = build_model(dev)
    eng = model.engine(BS, IMG, IMG, torch.float32)
    n_launch = _lib.lib().y3_model_num_launches(eng.handle)
    # two distinct resident input batches (157 MB each > 126 MB L2), alternated so no step re-reads a cached input
    xs = [torch.rand(BS, 3, IMG, IMG, device=dev, generator=torch.Generator(device=dev).manual_seed(1 + i)) for i in range(2)]
    graphs = [eng.capture(x) for x in xs]  # one graph per resident input: a step is exactly the 76 launches of a forward

    def step(i):
        graphs[i & 1].replay()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    eng.check_errors()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_load0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    clocks = None
    if rank == 0:
        # the clock samples must come from this workload: keep replaying it (untimed) until >= 0.6 s of load were observed
        i = args.steps
        while time.time() - t_load0 < 0.6:
            step(i)
            i += 1
            if i % 8 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        clocks = sampler.stop(t_load0, time.time())
        clocks["window_s"] = round(time.time() - t_load0, 3)
        clocks["timed_region_s"] = round(ms_total / 1e3, 3)
    ms_total = aggregate(ms_total, dev)
    if world > 1:
        dist.barrier()
    value = world * BS * args.steps / (ms_total / 1e3)

    # ---- e2e: host uint8 images -> H2D -> forward -> decode -> NMS -> D2H, through the public Pipeline
    pipe = Pipeline(model, BS, IMG, IMG, conf_thres=0.25, iou_thres=0.45, max_det=300)
    hosts = [torch.randint(0, 256, (BS, 3, IMG, IMG), dtype=torch.uint8, generator=torch.Generator().manual_seed(7 + i)).pin_memory()
             for i in range(2)]
    for i in range(3):
        pipe(hosts[i & 1])
    for _ in pipe.stream(hosts[i & 1] for i in range(3)):
        pass
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    n_det = 0
    for dets in pipe.stream(hosts[i & 1] for i in range(args.steps)):  # every step: H2D of its images, D2H of its boxes
        n_det += sum(d.shape[0] for d in dets)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    # the synchronous per-batch call (one batch in flight, as the reference's detect.py loop), for comparison
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        pipe(hosts[i & 1])
    torch.cuda.synchronize()
    ms_e2e_sync = (time.perf_counter() - t0) * 1e3
    ms_e2e = aggregate(ms_e2e, dev)
    e2e = world * BS * args.steps / (ms_e2e / 1e3)

    # ---- NMS sweep (BASELINE config 5): synthetic [bs,25200,85] fp32 resident in HBM, device pipeline only (no D2H)
    from yolov3_b200.nms import nms_batched

    from yolov3_b200.synth import synth_predictions

    pred = synth_predictions(BS, n_rows=25200, nc=80, seed=3).to(dev)
    nms_res = {}
    for conf, iou, ml in ((0.25, 0.45, False), (0.001, 0.6, False), (0.001, 0.6, True)):
        for _ in range(3):
            nms_batched(pred, conf, iou, multi_label=ml, max_det=300)
        torch.cuda.synchronize()
        e0.record()
        reps = 10
        for _ in range(reps):
            nms_batched(pred, conf, iou, multi_label=ml, max_det=300)
        e1.record()
        torch.cuda.synchronize()
        ms = aggregate(e0.elapsed_time(e1) / reps, dev)
        # HBM roofline of the whole NMS pipeline: ALGORITHMIC bytes = read z once, 8.568 MB/image
        gbs = BS * 25200 * 85 * 4 / (ms / 1e3) / 1e9
        nms_res[f"conf{conf}_iou{iou}_{'multi' if ml else 'single'}"] = {
            "input_boxes_per_s": world * BS * 25200 / (ms / 1e3), "ms_per_batch": ms,
            "hbm_gbs_per_gpu": gbs, "hbm_frac": gbs / peaks()["hbm"]}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (conv_tc): per-launch CUDA events on the launching stream
    from yolov3_b200.profile import time_ops

    per_op = time_ops(eng, xs[0], iters=max(3, min(10, args.steps)))
    conv_ms = sum(o["ms"] for o in per_op if o["kind"] == "conv_tc")
    all_ms = sum(o["ms"] for o in per_op)
    conv_tflop = (GFLOP_PER_IMG - GFLOP_LAYER0) * BS / 1e3
    pk = peaks()
    achieved = conv_tflop / (conv_ms / 1e3)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["tf_sustained"], "traffic": None,
                "kernel": "conv_tc_kernel (74 launches/step)", "kernel_ms_per_step": conv_ms,
                "kernel_share_of_step": conv_ms / all_ms, "peak_source": pk["source"] + " sustained bf16 (MEASURED_PEAKS.json)",
                # `traffic` stays null: `achieved` aggregates 74 launches of 23 different shapes.  One ncu --set full capture
                # of the largest layer group (recorded, not re-measured here): dram read+write vs algorithmic in+res+out+w
                "traffic_sample": {"kernel": "conv_tc 128->256 3x3 s1 @80x80 bs32 +res (8 of the 74 launches)",
                                   "dram_bytes_per_launch": 226.9e6, "algorithmic_bytes_per_launch": 262.7e6,
                                   "source": "profiles/r01_ncu_conv_tc_final_summary.txt"}}
    dec = [o for o in per_op if o["kind"] == "decode"]
    if dec:  # Detect decode: read the head logits + write z = 17.1 MB/image algorithmic
        roofline["decode_hbm"] = {"bound": "hbm", "achieved": dec[0]["gbs"], "peak": pk["hbm"], "unit": "GB/s",
                                  "frac": dec[0]["gbs"] / pk["hbm"], "ms": dec[0]["ms"]}
    if args.per_op:
        Path(args.per_op).parent.mkdir(parents=True, exist_ok=True)
        Path(args.per_op).write_text(json.dumps(per_op, indent=1))

    cpu = None
    if not args.no_cpu_baseline:
        n_img = 4
        rate, cores = cpu_forward_rate(n_img, 2, warmup=1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_img} of the {BS} images per step, 2 timed passes after 1 warm-up (oracle port, torch CPU fp32)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"yolov3.yaml forward+decode, {IMG}x{IMG}, bs {BS}/GPU, random-init weights, folded BN",
                   "imgsz": IMG, "batch_per_gpu": BS, "global_batch": BS * world, "parallelism": f"replicas x{world} (no collective)",
                   "l2": "inputs larger than L2: two 157 MB fp32 batches alternated; activations 6 GB/step",
                   "cuda_graph": True},
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                "ms_per_step": ms_e2e / args.steps, "path": "Pipeline.stream: uint8 H2D -> forward -> decode -> NMS(0.25/0.45/300) -> D2H, two batches in flight",
                "sync_call_ms_per_step": ms_e2e_sync / args.steps},
        "gpu_launches": n_launch * args.steps,
        "clocks": clocks,
        "nms": {"workload": f"synthetic [bs {BS}/GPU, 25200, 85] fp32 (SURVEY §8d config 5), max_det 300, device-resident, "
                            "sync-free y3_nms_batched", "unit": "input boxes/s", **nms_res},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
