#!/usr/bin/env python
"""bench.py — Flux-VAE HDR decode throughput on B200 (BASELINE.json metric: megapixels/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path (HDRVAEDecode.simple_hdr_decode's happy path, reference
hdr_vae_decode.py:62-195, ONE decoder pass) over one batch of synthetic latents.
Workload per GPU: BASELINE config C2 — 4x16x128x128 latents -> 4 x 1024x1024, "moderate" mode, bf16
tensor-core decoder (fp16 operands, the 16-bit mode that meets the 1e-2 parity tolerance).  N > 1: batch sharding (config C3 style): every rank decodes 4 more images and
the batch-global HDR statistics are all-reduced over NCCL between epilogue phase A and B ("weak").

value : device-timed MP/s, latents resident in HBM, max over ranks.
e2e   : the same metric through the node API with HOST buffers (pinned latent H2D + IMAGE D2H inside
        the timed region).
roofline : the tcgen05 implicit-GEMM conv kernel (dominant): algorithmic conv FLOPs of one step
        (9.4628 MFLOP per output pixel) / summed device time of its launches in one
        step (CUDA events on the launching stream), against the measured sustained bf16 peak.
cpu_baseline : the oracle port of the reference node on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "flux_vae_hdr_decode_megapixels_per_s"
UNIT = "MP/s"
CONV_MFLOP_PER_PX = 9.4628          # 3x3/1x1 convs of the decoder, resolution independent
PER_GPU_BATCH, LATENT = 4, 128      # config C2
MODE = "moderate"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1400.0, 6650.0, "fallback"      # B200_PROFILING.md fallback (sustained bf16, HBM copy)


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled DURING the timed region (B200_PROFILING.md recipe).
    NVML in-process (microseconds per query, so even a 0.2 s region yields dozens of samples on an 8-GPU box where
    spawning nvidia-smi takes longer than the region); `nvidia-smi -lms` as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.rows, self.proc, self.index, self.stop, self.t, self.src = [], None, index, False, None, None

    def _nvml_loop(self, nv, h):
        bits = [(getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                (getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                (getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                (getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = 0
        while not self.stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b, _ in bits])
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.src = "nvml"
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.t.start()
            return self
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.src = "nvidia-smi"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        self.stop = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        elif self.t is not None:
            self.t.join(timeout=1)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(self.NAMES, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": self.src}


def cpu_reference_run(steps: int, warmup: int, sample_latent: int = 32):
    """The reference node's CPU path (oracle port: fp32 PyTorch eager decoder + the restated HDR math,
    with the reference's real call structure: TWO decoder passes + a third conv_out per call,
    hdr_vae_decode.py:859,876,1022) on all host cores.  Bounded sample: one 1x16xSxS latent per step."""
    import torch
    from oracle import hdr_oracle as ho
    from oracle.flux_decoder import build_decoder, make_latent
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dec = build_decoder(0)
    z = make_latent(1, sample_latent, sample_latent, seed=1234)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            _ = dec(z)                                            # analyze_conv_out's vae.decode (:859)
            out, st, _pre = ho.simple_hdr_decode(dec, z, MODE, 1.0)   # second decode (:1022) + conv_out + HDR math
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    mp = (8 * sample_latent) ** 2 / 1e6
    per = sum(times) / len(times)
    return {"value": mp / per, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} x (1x16x{sample_latent}x{sample_latent} latent -> {8 * sample_latent}^2, {MODE}; "
                      f"2 decoder passes + conv_out as the reference node does), {per:.2f} s per call, fp32 torch CPU"}, per


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, per = cpu_reference_run(max(1, min(args.steps, 3)), 1)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, min(args.steps, 3)), "warmup": 1, "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C2: {PER_GPU_BATCH}x16x{LATENT}x{LATENT} latents -> 1024x1024, {MODE} "
                                   "(each reference step is a bounded sample of it, see cpu_baseline.sample)"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def guard_stdout() -> None:
    """stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner with printf when
    NCCL_DEBUG is set on the box, regardless of NCCL_DEBUG_FILE), so file descriptor 1 is pointed at stderr for the whole
    run and the JSON line goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-aux", action="store_true", help="skip the auxiliary 4096^2 single-GPU measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from vae_decode_hdr_b200 import NODE_CLASS_MAPPINGS, _native
    from vae_decode_hdr_b200.engine import HdrVaeEngine
    from vae_decode_hdr_b200.sharding import decode_batch_sharded
    from vae_decode_hdr_b200.synthetic import SyntheticVAE, random_decoder_state_dict, synthetic_latent

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL writes its debug output (the "NCCL version ..." banner on boxes that set NCCL_DEBUG) to stdout by
        # default: send it to stderr so that stdout stays the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    lib = _native.load_library()

    B, L = PER_GPU_BATCH, LATENT
    sd = random_decoder_state_dict(0)
    engine = HdrVaeEngine(sd, dev)
    z_dev = synthetic_latent(B, L, L, seed=1234 + rank).to(dev)
    mp_per_rank = B * (8 * L) ** 2 / 1e6

    def step_device():
        if world > 1:
            return decode_batch_sharded(engine, z_dev, MODE, 1.0, want_stats=False)
        return engine.decode(z_dev, MODE, 1.0, want_stats=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step_device()
    barrier()
    n0 = lib.hdrvae_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step_device()
        e1.record()
        barrier()
    launches = lib.hdrvae_launch_count() - n0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_step = ms_total / args.steps
    value = mp_per_rank * world / (ms_step / 1e3)

    # ---- roofline of the dominant kernel (tcgen05 conv): per-op CUDA events over one extra step
    peak_tf, peak_gbs, peak_src = _peaks()
    prof_path = os.path.join(tempfile.gettempdir(), f"hdrvae_prof_{os.getpid()}.tsv")
    lib.hdrvae_profile_begin()
    engine.decode(z_dev, MODE, 1.0, want_stats=False)
    lib.hdrvae_profile_end(prof_path.encode())
    conv_ms = gn_ms = attn_ms = epi_ms = 0.0
    n_conv = 0
    with open(prof_path) as f:
        for ln in f:
            name, t = ln.split("\t")[0].strip(), float(ln.split("\t")[1].split()[0])
            if name.startswith("conv3x3 128->8 "):
                continue      # conv_out on the tensor cores: nested inside (and counted with) "epilogue phase A"
            if name.startswith("conv"):
                conv_ms += t; n_conv += 1
            elif name.startswith("groupnorm"):
                gn_ms += t
            elif name.startswith("attention"):
                attn_ms += t
            elif name.startswith("epilogue"):
                epi_ms += t
    os.unlink(prof_path)
    conv_flops = CONV_MFLOP_PER_PX * 1e6 * B * (8 * L) ** 2 - 2.0 * 1152 * 3 * B * (8 * L) ** 2  # conv_out runs in the epilogue
    achieved_tf = conv_flops / (conv_ms / 1e3) / 1e12
    # the three upsample convs run as four 2x2-tap phase convs (nearest-2x folded into the load): 4/9 of their
    # algorithmic FLOPs are executed.  Per output pixel of the decoder: up convs = 2*9*(512*512/16 + 512*512/4 + 256*256)
    up_flops = 2.0 * 9 * (512 * 512 / 16.0 + 512 * 512 / 4.0 + 256 * 256) * B * (8 * L) ** 2
    executed_tf = (conv_flops - up_flops * 5.0 / 9.0) / (conv_ms / 1e3) / 1e12
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 implicit-GEMM conv, all conv layers of one step)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "peak_source": f"{peak_src} bf16_tflops_sustained",
                # ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the 61 gemm_tc launches of one C2 step
                # (profiles/r01_step9_ncu_launch_summary_c2.json, captured on this workload with the final build: 37.1 GB
                # read + 40.3 GB written; includes the attention GEMMs and conv_out).  The kernel is tensor bound; the
                # figure shows there is no re-read waste (GroupNorm: 43.2 GB measured vs 44.1 GB algorithmic)
                "traffic": (77.41e9 if (B, L) == (4, 128) else None), "traffic_unit": "bytes per step, all conv launches (ncu)",
                "achieved_executed": executed_tf, "frac_executed": executed_tf / peak_tf,
                "note": "achieved = algorithmic conv FLOPs (SURVEY 8d) / summed conv launch time of one step; "
                        "achieved_executed discounts the 5/9 of the upsample convs' FLOPs that phase decomposition removes",
                "step_breakdown_ms": {"conv": conv_ms, "groupnorm_silu": gn_ms, "attention": attn_ms, "epilogue": epi_ms},
                "groupnorm_gbs": (1837.1e6 * B * 6.0 / (gn_ms / 1e3) / 1e9) if gn_ms > 0 else None,
                "hbm_peak_gbs": peak_gbs}

    # ---- e2e through the node API with host buffers
    vae = SyntheticVAE(sd, device=dev, output_device="cpu")
    node = NODE_CLASS_MAPPINGS["HDRVAEDecode"]()
    node.adopt_engine(vae, dev, engine)          # reuse the packed weights (same state dict)
    z_host = synthetic_latent(B, L, L, seed=1234 + rank).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        (img,) = node.simple_hdr_decode({"samples": z_host}, vae, hdr_mode=MODE)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        (img,) = node.simple_hdr_decode({"samples": z_host}, vae, hdr_mode=MODE)
        assert img.device.type == "cpu"
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = {"value": mp_per_rank * world / (float(dt.item()) / e2e_steps), "unit": UNIT,
           "h2d_bytes_per_step": z_host.numel() * 4, "d2h_bytes_per_step": img.numel() * 4,
           "note": "node API, pinned host latent in, host IMAGE out; N>1: independent per-rank node calls"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16", "data": "synthetic",
            "config": {"workload": f"C2 per GPU: {B}x16x{L}x{L} random latents -> {B} x {8 * L}x{8 * L}, mode {MODE} "
                                   "(smart expansion x3), Flux.1 AE decoder random-init, fp16 operands / fp32 accumulate / fp32 residual stream",
                       "precision": "fp16 tensor-core operands (16-bit, same width and rate as the bf16 BASELINE names; bf16 "
                                    "misses its 1e-2 parity bar, DESIGN.md), fp32 accumulate, fp32 residual stream",
                       "global_batch": B * world,
                       "parallelism": "single GPU" if world == 1 else f"batch-sharded dp{world} + all-reduce of HDR statistics",
                       "l2": "inputs larger than L2: ~6 GB of activations stream through HBM every step (L2 126 MB)"},
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary()}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_reference_run(2, 1)
    if world == 1 and not args.no_aux:
        # BASELINE.json quotes the metric at 1024^2 AND 4096^2: config C4 (1x16x512x512 -> 4096^2, "aggressive") on this
        # one GPU, outside the timed region of the headline value; the 8-GPU row-tiled number is in profiles/ (58.4 ms).
        try:
            engine._workspace = None
            torch.cuda.empty_cache()
            z4 = synthetic_latent(1, 512, 512, seed=1234).to(dev)
            engine.decode(z4, "aggressive", want_stats=False)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                out4, _ = engine.decode(z4, "aggressive", want_stats=False)
            e1.record()
            torch.cuda.synchronize(dev)
            ms4 = e0.elapsed_time(e1) / 2
            line["aux_c4_4096"] = {"workload": "C4 on one GPU: 1x16x512x512 latent -> 4096x4096, aggressive (not row-tiled)",
                                   "ms_per_image": ms4, "value": 16.777216 / (ms4 / 1e3), "unit": UNIT,
                                   "finite": bool(torch.isfinite(out4).all())}
        except Exception as exc:      # never lose the headline line to the auxiliary measurement
            line["aux_c4_4096"] = {"error": repr(exc)[:200]}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
