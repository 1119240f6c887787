#!/usr/bin/env python
"""bench.py — attack iterations/s and clip-frames/s of the I3D flickering-attack loop.

  python bench.py [--gpus N] [--steps K] [--warmup W]          engine arm (libfav, sm_100a)
  python bench.py --impl reference [...]                        reference arm: the CPU restatement of
                                                                the reference path on the host cores
For N > 1 launch with torchrun (one rank per GPU, NCCL); RANK/LOCAL_RANK/WORLD_SIZE come from the env.

Workload at N=1 = BASELINE.json configs[1]: I3D single-class-generalisation flickering attack,
batch 8 x 64x224x224x3 synthetic uint8 clips, random-init weights.  For N > 1 every rank keeps the
same per-GPU batch (weak scaling, global batch 8N, one sum-all-reduce of the [T,3]+scalars buffer per
step).  One "step" = apply delta -> forward -> loss -> backward to delta -> all-reduce -> Adam.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "clip_frames_per_sec"
UNIT = "clip-frames/s"


def i3d_forward_macs(frames, height=224, width=224):
    """Forward conv MACs per clip from the layer table (i3d.py:168-474)."""
    from flickering_adversarial_video_b200.synthetic import i3d_units

    def same(n, s):
        return -(-n // s)
    t1, h1, w1 = same(frames, 2), same(height, 2), same(width, 2)
    h2, w2 = same(h1, 2), same(w1, 2)
    h3, w3 = same(h2, 2), same(w2, 2)
    t4, h4, w4 = same(t1, 2), same(h3, 2), same(w3, 2)
    t5, h5, w5 = same(t4, 2), same(h4, 2), same(w4, 2)
    macs = 0
    for scope, k, cin, cout in i3d_units():
        if scope == "Conv3d_1a_7x7":
            pos = t1 * h1 * w1
        elif scope.startswith("Conv3d_2"):
            pos = t1 * h2 * w2
        elif scope.startswith("Mixed_3"):
            pos = t1 * h3 * w3
        elif scope.startswith("Mixed_4"):
            pos = t4 * h4 * w4
        else:
            pos = t5 * h5 * w5
        macs += pos * k ** 3 * cin * cout
    macs += (t5 - 1) * 1024 * 400
    return macs


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", 1391.3)), "hbm": float(d.get("hbm_gbs", 6548.2)),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi style clock/throttle sampling DURING the timed region (pynvml)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_reference_step_time(frames, steps, warmup, threads):
    """Time the oracle's attack step (CPU restatement of the reference path) on one clip."""
    import torch
    from flickering_adversarial_video_b200 import synthetic
    from oracle import oracle_i3d
    torch.set_num_threads(threads)
    weights = synthetic.i3d_weights(seed=0)
    model = oracle_i3d.OracleI3D(weights)
    clip = synthetic.clips_u8(1, frames, seed=1000)
    x = oracle_i3d.normalize_u8(clip)
    labels = torch.zeros(1, dtype=torch.int64)
    delta = torch.zeros((frames, 3))
    opt = oracle_i3d.TFAdam((frames, 3))
    cfg = dict(improve_loss=True, margin=0.05, beta0=10.0, beta1=0.5, beta2=0.5, beta3=0.5, lr=1e-3)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = oracle_i3d.attack_step(model, x, labels, delta, cfg, opt=opt)
        delta = out["delta_new"]
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = len(os.sched_getaffinity(0))
    sec = cpu_reference_step_time(args.frames, max(1, args.steps), max(1, min(args.warmup, 2)), threads)
    value = args.frames / sec     # one clip per step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "iters_per_sec": 1.0 / sec,
        "config": {"workload": f"I3D class-generalisation flickering attack, {args.frames}x224x224x3 clips "
                               f"(BASELINE.json configs[1]); reference arm = CPU restatement of the TF1.15 path "
                               f"(TF1.15 not installable: SURVEY D5), one clip per step",
                   "batch_per_step": 1, "frames": args.frames},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} attack iterations (fwd + bwd-to-input + Adam) on 1 clip of "
                                   f"{args.frames} frames, torch CPU fp32, {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_engine(args):
    import torch
    import torch.distributed as dist
    from flickering_adversarial_video_b200 import synthetic, _lib
    from flickering_adversarial_video_b200.attack import FlickerAttack

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    B, T = args.batch, args.frames
    cfg = {"IMPROVE_ADV_LOSS": True, "TARGETED_ATTACK": False, "USE_LOGITS": False, "PROB_MARGIN": 0.05,
           "LAMBDA": 10.0, "BETA_1": 0.5, "BETA_2": 0.5}     # run_config.yml CLASS_GEN_ATTACK
    weights = synthetic.i3d_weights(seed=0)
    atk = FlickerAttack(weights, B, T, cfg, device=local_rank)
    lib = _lib.load()

    # resident synthetic inputs: `pool` different batches per rank (each step sees a different batch)
    pool = args.pool
    clips = [synthetic.clips_u8(B, T, seed=1000 + rank * 100 + i, device=dev) for i in range(pool)]
    labels = []
    for c in clips:
        labels.append(atk.predict(c, adv_flag=0.0).argmax(-1))   # reference attacks correctly classified clips
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    use_graph = not args.no_graph
    l0 = lib.fav_launch_count()
    atk.step(clips[0], labels[0])
    launches_per_step = lib.fav_launch_count() - l0       # kernels one step launches (a graph replays the same nodes)
    if use_graph:
        # one CUDA graph per resident batch (the graph is bound to the batch's device buffers)
        graphs = []
        for i in range(pool):
            graphs.append(atk.capture(clips[i], labels[i]))
        atk.reset()
        step_fn = lambda i: graphs[i % pool].replay()
    else:
        step_fn = lambda i: atk.step(clips[i % pool], labels[i % pool])
    for i in range(args.warmup):
        step_fn(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.fav_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step_fn(i)
    ev1.record()
    barrier()
    launches = launches_per_step * args.steps if use_graph else lib.fav_launch_count() - launches0
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    iters_per_sec = 1e3 / ms_per_step
    value = iters_per_sec * B * world * T

    # ---------------- end-to-end: pinned host clips in, host scalars out ----------------
    host_clips = [c.cpu().pin_memory() for c in clips[:2]]
    host_labels = [l.cpu().pin_memory() for l in labels[:2]]
    for i in range(2):   # warm the staging path
        slot = atk.prefetch(host_clips[i % 2], host_labels[i % 2])
        atk.step_staged(slot)
    if use_graph:
        atk.capture_staged()   # the public API's own graph mode: one CUDA graph per staging slot
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    slot = atk.prefetch(host_clips[0], host_labels[0])
    pending = None
    losses = []
    for i in range(args.steps):
        nxt = atk.prefetch(host_clips[(i + 1) % 2], host_labels[(i + 1) % 2]) if i + 1 < args.steps else None
        out = atk.step_staged(slot)
        if pending is not None:
            pending[1].synchronize()
            losses.append(float(pending[0][_lib.S_TOTAL_LOSS]))
        pending = out
        slot = nxt
    pending[1].synchronize()
    losses.append(float(pending[0][_lib.S_TOTAL_LOSS]))
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = args.steps * B * world * T / (ms_e2e * 1e-3)
    h2d = host_clips[0].numel() * host_clips[0].element_size() + host_labels[0].numel() * 8
    d2h = _lib.S_COUNT * 4

    # ---------------- per-kernel-family timing, live, CUDA events on the launch stream ----------------
    import ctypes
    prof_steps = 3
    lib.fav_profile_begin()
    for i in range(prof_steps):
        atk.step(clips[i % pool], labels[i % pool])
    buf = (ctypes.c_double * (4 * len(_lib.PROF_KINDS)))()
    lib.fav_profile_end(buf, len(buf))
    peaks = load_peaks()
    kernels = {}
    for k, name in enumerate(_lib.PROF_KINDS):
        kms, n, fl, by = buf[4 * k], buf[4 * k + 1], buf[4 * k + 2], buf[4 * k + 3]
        if n == 0:
            continue
        ent = {"ms_per_step": kms / prof_steps, "launches_per_step": n / prof_steps, "avg_launch_us": 1e3 * kms / n}
        if fl > 0:
            ent["tflops"] = fl / kms / 1e9
            ent["frac_of_tensor_peak"] = ent["tflops"] / peaks["tflops"]
        if by > 0:
            ent["gbs"] = by / kms / 1e6
            ent["frac_of_hbm_peak"] = ent["gbs"] / peaks["hbm"]
        kernels[name] = ent

    # ---------------- roofline: the dominant kernel (conv_halo_kernel) + the whole step ----------------
    flop_per_clip_iter = 4.0 * i3d_forward_macs(T)
    step_tflops = iters_per_sec * B * flop_per_clip_iter / 1e12      # per GPU, whole step
    dom = max((k for k in kernels if "tflops" in kernels[k]), key=lambda k: kernels[k]["ms_per_step"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom, {}).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": dom, "achieved": kernels[dom]["tflops"], "peak": peaks["tflops"],
                "unit": "TFLOP/s", "frac": kernels[dom]["tflops"] / peaks["tflops"], "traffic": traffic,
                "avg_launch_us": kernels[dom]["avg_launch_us"], "launches_per_step": kernels[dom]["launches_per_step"],
                "share_of_step": kernels[dom]["ms_per_step"] / sum(v["ms_per_step"] for v in kernels.values()),
                "step": {"achieved": step_tflops, "frac": step_tflops / peaks["tflops"]},
                "note": f"kernel: algorithmic FLOPs (2 x MACs, real channels) of its launches / their CUDA-event time, "
                        f"{prof_steps} un-graphed steps; step: 4 x forward conv MACs = {flop_per_clip_iter / 1e9:.2f} "
                        f"GFLOP per clip-iteration x {B} clips over the whole step time; peak = {peaks['source']}; "
                        f"traffic = ncu dram bytes per launch (profiles/r01_traffic.json)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "iters_per_sec": iters_per_sec,
        "config": {"workload": f"I3D class-generalisation flickering attack (BASELINE.json configs[1]), "
                               f"{B} x {T}x224x224x3 uint8 clips per GPU, random-init weights",
                   "batch_per_gpu": B, "global_batch": B * world, "frames": T, "resident_batches": pool,
                   "cuda_graph": use_graph,
                   "l2": f"each step touches ~{atk.eng.device_bytes / 2**30:.1f} GiB of activations/gradients "
                         f"(>> 126 MB L2) and a different clip batch"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "last_total_loss": losses[-1]},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernels": kernels,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0))
        sec = cpu_reference_step_time(T, 3, 1, threads)
        line["cpu_baseline"] = {"value": T / sec, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"3 attack iterations on 1 clip of {T} frames (oracle/oracle_i3d.py, "
                                          f"torch CPU fp32, {threads} threads), {sec:.2f} s/iter"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        # all ranks are past their last collective; leave without the NCCL / CUDA-graph teardown,
        # which can block on process-group destruction when captured collectives are still referenced
        barrier()
        os._exit(0)
    atk.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="clips per GPU per step")
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--pool", type=int, default=3, help="resident clip batches per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels one by one instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "engine":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_engine(args)


if __name__ == "__main__":
    sys.exit(main())
