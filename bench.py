#!/usr/bin/env python
"""bench.py — reference-grid views/sec on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one synthetic reference grid (BASELINE config C2/C3):
    K1  render 16 views 512x512, flat 128 samples/ray through the nerfacto field   (sgn_render_views)
    K2/K3  AABB mask + elliptical dilation + depth condition                         (sgn_mask_condition)
    K4  paste the tiles into the 2048x2048 image / mask / condition sheets           (sgn_sheet_paste)
    K5-K9 one ControlNet-depth SDXL UNet step (CFG batch 2) on the sheet latent      (sgn_unet_*; when built)

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's Python path (oracle port) on host cores

Multi-GPU (launched by torchrun, one rank per GPU): weak scaling.  A step produces N grids; every rank renders
a per-view shard (views v with v % N == rank) of EVERY grid, one NCCL all-gather assembles the tiles, and rank r
pastes / denoises grid r.  Per-rank work is exactly one metric unit for every N.

Timing: per-step CUDA events on the launching stream, L2 flushed (256 MiB memset) between steps outside the
events, barrier + synchronize on both sides of the K timed steps, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# NCCL announces its version on STDOUT at NCCL_DEBUG=VERSION; the contract is ONE JSON line on stdout
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# ---------------------------------------------------------------------------------------------- workload
ROWS = COLS = 4
VIEWS = ROWS * COLS
H = W = 512
SAMPLES = 128
LEVELS, CORNERS, FEATS = 16, 8, 2
RENDER_BYTES_PER_SAMPLE = LEVELS * CORNERS * FEATS * 4      # 1024 B of hash-table gathers
RENDER_BYTES_PER_RAY = 52                                   # o, d, near, far in; rgb, depth, acc out
UNET_TFLOP_PER_STEP = 103.96                                # UNet 35.91 + ControlNet 16.08 per sample, x2 CFG

METRIC = "reference-grid views/sec (render+1 UNet step) @4x4x512^2"
UNIT = "grids/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tensor=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured")
    return dict(hbm=6650.0, tensor=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi sampler running DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- reference arm
def cpu_oracle_sample(rays_per_sample: int):
    """Time the oracle (literal torch restatement of nerfstudio's torch nerfacto eval = the reference's Python
    path) on a bounded sample of the C2 workload: `rays_per_sample` rays of view 0 x 128 samples, in chunks of
    32 768 rays as `eval_num_rays_per_chunk` (signerf_config.py:32)."""
    from oracle import nerfacto_ref as R
    from tests.helpers import ring_cameras
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = R.make_model(0)
    c2w, intr = ring_cameras(VIEWS, W, H)
    rays = R.generate_rays(c2w[0], *intr[0].tolist(), W, H)
    o, d = rays.origins[:rays_per_sample], rays.directions[:rays_per_sample]

    def run():
        t = time.perf_counter()
        for i in range(0, o.shape[0], 1 << 15):
            R.render_rays(m, o[i:i + (1 << 15)], d[i:i + (1 << 15)], "flat", SAMPLES)
        return time.perf_counter() - t

    return run, cores, rays_per_sample


def unet_tflop(sheet: int) -> float:
    """Algorithmic TFLOP of one UNet + ControlNet step with CFG (batch 2) on a sheet x sheet image: everything but
    self-attention scales with the pixel count, self-attention with its square (34.7 of 103.96 TFLOP at
    2048^2 are QK^T / PV)."""
    r = (sheet / 2048.0) ** 2
    return (UNET_TFLOP_PER_STEP - 34.7) * r + 34.7 * r * r


def cpu_unet_sample(sheet: int):
    """The reference's diffusion arithmetic (fp32 torch, `--no-half`) = oracle/sdxl_ref.py at full SDXL width on the host
    cores, one CFG step incl. ControlNet on a `sheet`^2 image; scaled to the 2048^2 sheet by algorithmic FLOPs
    (1024^2 measured, x5.3)."""
    from oracle import sdxl_ref as X
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = X.UNetConfig()
    unet, ctrl = X.make_models(cfg, fast_init=True)
    h = sheet // 8
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 4, h, h, generator=g)
    ctx, y = torch.randn(2, 77, cfg.context_dim, generator=g), torch.randn(2, cfg.adm_in_channels, generator=g)
    hint, noise = torch.rand(1, 3, 8 * h, 8 * h, generator=g), torch.randn(1, 4, h, h, generator=g)

    def run():
        t = time.perf_counter()
        X.denoise_step(unet, ctrl, x, 13.0, 10.0, ctx, y, hint, noise)
        return time.perf_counter() - t

    return run, cores, unet_tflop(2048) / unet_tflop(sheet)


def cpu_baseline(rays: int, render_runs: int, unet_sheet: int, unet_runs: int, with_unet: bool = True):
    """-> (grids/s, cores, sample description, parts) of the CPU path on bounded samples of the benchmark workload."""
    run, cores, n = cpu_oracle_sample(rays)
    if n < H * W:
        run()                                   # warm-up (a full view is its own warm-up: ~1 minute of CPU time)
    t_r = float(np.mean([run() for _ in range(render_runs)]))
    per_grid = t_r / n * VIEWS * H * W
    parts = {"render_sample_rays": n, "render_sample_s": t_r, "render_s_per_grid_scaled": per_grid}
    what = "one full 512x512 view" if n == H * W else f"{n} rays of view 0"
    sample = (f"render: {what} x {SAMPLES} samples through oracle/nerfacto_ref.py in {t_r:.1f} s measured, x{VIEWS * H * W / n:.0f} "
              f"to 16 views = {per_grid:.1f} s/grid")
    if with_unet:
        urun, _, scale = cpu_unet_sample(unet_sheet)
        t = float(np.mean([urun() for _ in range(unet_runs)]))
        per_grid += t * scale
        parts.update({"unet_sample_sheet": unet_sheet, "unet_sample_s": t, "unet_flop_scale": scale, "unet_s_per_grid_scaled": t * scale})
        sample += (f"; diffusion: one fp32 UNet+ControlNet CFG step of oracle/sdxl_ref.py (full SDXL width) on a "
                   f"{unet_sheet}^2 sheet = {t:.1f} s measured, x{scale:.2f} by algorithmic FLOPs to the 2048^2 sheet")
    return 1.0 / per_grid, cores, sample, parts


def reference_arm(args):
    """The reference's own CPU implementation of the path = the oracle ports (nothing of the reference is installable
    here: SURVEY §0), on all host cores; rank 0 only.  Samples as SURVEY §8d specifies them: ONE full 512 x 512 x 128
    view (1/16 of the grid's render) and ONE UNet+ControlNet step on a 1024^2 sheet (19.5 of the 103.96 TFLOP), each
    measured once per run (they take about a minute each on 16 cores) and scaled to the grid."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    val, cores, sample, parts = cpu_baseline(rays=H * W, render_runs=1, unet_sheet=1024, unet_runs=1, with_unet=not args.no_unet)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3 4x4 grid 512x512, flat 128 samples/ray + 1 SDXL+ControlNet UNet step (CFG 2) on the "
                                   "2048^2 sheet; CPU arm measured on bounded samples and scaled (see cpu_baseline.sample)",
                       "measured_vs_scaled": parts, "wall_s": None},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["config"]["wall_s"] = round(time.perf_counter() - t0, 1)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- CUDA arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-unet", action="store_true", help="time the render half only (profiling)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end and per-kernel passes (ncu runs only)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU tile exchange: fused NVLink peer stores (default) or one NCCL all-gather")
    ap.add_argument("--config5-cameras", type=int, default=1,
                    help="dataset cameras PER RANK of the BASELINE-config-5 generation sample (0 = skip)")
    ap.add_argument("--config5-total", action="store_true", help="--config5-cameras is the TOTAL camera count (30 = the named job)")
    ap.add_argument("--config5-rounds", type=int, default=1, help="generation + fine-tune rounds (BASELINE config 5 names 3)")
    ap.add_argument("--config5-train-steps", type=int, default=20, help="fine-tune steps after each generation (0 = none)")
    ap.add_argument("--ncu-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps (run under `ncu --profile-from-start off`)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from signerf_b200 import _lib, ops, sharding, synthetic
    from signerf_b200.sheet import ReferenceSheetRenderer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: signerf_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # communicator creation prints an "NCCL version" banner on STDOUT; the contract is ONE JSON line there, so fd 1
        # points at stderr while the process group and its first collective come up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    N = world

    for opt, env in (("gemm_pair_stages", "SGN_GEMM_PAIR_STAGES"), ("render_ctas_per_sm", "SGN_RENDER_CTAS_PER_SM")):
        if os.environ.get(env):   # tuning knobs of the library (sgn_set_option), for A/B runs
            _lib.set_option(opt, int(os.environ[env]))
    fld = synthetic.random_field(seed=0, device=dev, dense=True, with_proposals=False)
    layout = ops.SheetLayout(ROWS, COLS, H, W, 0)
    ropts = ops.RenderOptions(mode="flat", num_samples=SAMPLES)
    mopts = ops.MaskOptions()
    sheet = ReferenceSheetRenderer(fld, layout, H, W, ropts, mopts)

    # cameras of the N grids of one step: grid g is the benchmark ring rotated by g * 7 degrees
    c2w_all, intr_all = [], []
    for g in range(N):
        c, i = synthetic.camera_ring(VIEWS, W, H, phi_deg=(7.0 * g, 300.0 + 7.0 * g))
        c2w_all.append(c)
        intr_all.append(i)
    c2w_all = torch.stack(c2w_all)      # [N,16,3,4]
    intr_all = torch.stack(intr_all)    # [N,16,4]
    # per-view shard of every grid for this rank
    mine = sharding.views_of_rank(VIEWS, N, rank)
    c2w_h = c2w_all[:, mine].reshape(-1, 3, 4).contiguous().pin_memory()
    intr_h = intr_all[:, mine].reshape(-1, 4).contiguous().pin_memory()
    c2w_d, intr_d = c2w_h.to(dev), intr_h.to(dev)
    nv = c2w_d.shape[0]                 # == 16 for every N
    assert nv == VIEWS

    unet = None
    if not args.no_unet:
        from signerf_b200 import unet as unet_mod
        unet = unet_mod.BenchUNet(dev, sheet_hw=(layout.height, layout.width), seed=0)

    tiles = torch.empty((N, VIEWS, H, W, 6), dtype=torch.float32, device=dev) if N > 1 else None
    # tile exchange: peer stores over NVLink fused with the packing (symmetric memory) - NCCL all-gather when the
    # process group cannot provide symmetric memory (--exchange nccl forces it)
    exchange, exchange_name = None, "single GPU"
    if N > 1:
        exchange_name = f"per-view x{N} + NCCL all-gather"
        if args.exchange == "peer":
            try:
                exchange = sharding.PeerTileExchange(N, rank, VIEWS, H, W, dev)
                exchange_name = f"per-view x{N} + fused NVLink peer stores (symmetric memory)"
            except Exception as e:  # noqa: BLE001
                print(f"[bench] symmetric-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL all-gather",
                      file=sys.stderr)
        ok = torch.tensor([1 if exchange is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)      # all ranks must agree on the path
        if int(ok) == 0:
            exchange, exchange_name = None, f"per-view x{N} + NCCL all-gather"
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    k1_events, unet_events = [], []

    def step(c2w, intr, time_k1=False):
        if time_k1:
            a, b = ev(), ev()
            a.record()
        rgb, depth = ops.render_views(fld, c2w, intr, H, W, ropts)
        if time_k1:
            b.record()
            k1_events.append((a, b))
        mask, cond, _ = ops.mask_condition(c2w, intr, depth, mopts)
        if N > 1 and exchange is not None:
            rgb, _, cond, mask = sharding.unpack_tiles(exchange.exchange(rgb, depth, cond, mask))
        elif N > 1:
            # pack [rgb3 | depth | cond | mask] per pixel and all-gather the per-view shards of all N grids
            packed = sharding.pack_tiles(rgb, depth, cond, mask).view(N, len(mine), H, W, sharding.PACK_CHANNELS)
            sharding.gather_grids(packed, VIEWS, N, rank, out=tiles)
            rgb, _, cond, mask = sharding.unpack_tiles(tiles[rank])
        b_ = sheet.paste(rgb, mask, cond, 0)
        out = b_.image
        if unet is not None:
            if time_k1:
                a, b = ev(), ev()
                a.record()
            out = unet.step(b_.image, b_.mask, b_.condition)
            if time_k1:
                b.record()
                unet_events.append((a, b))
        return out

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(c2w_d, intr_d)
        flush.zero_()
    sync_all()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    evs = []
    sync_all()
    if args.ncu_range:
        torch.cuda.cudart().cudaProfilerStart()
    for _ in range(args.steps):
        a, b = ev(), ev()
        a.record()
        step(c2w_d, intr_d, time_k1=True)
        b.record()
        evs.append((a, b))
        flush.zero_()
    sync_all()
    if args.ncu_range:
        torch.cuda.cudart().cudaProfilerStop()
    launches = _lib.launch_count() - launches0
    if unet is not None:   # graph replays re-issue the launches counted while capturing
        launches += args.steps * unet.launches_per_step
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    k1_ms = sum(a.elapsed_time(b) for a, b in k1_events) / len(k1_events)
    unet_ms = sum(a.elapsed_time(b) for a, b in unet_events) / len(unet_events) if unet_events else 0.0

    # correctness carried by the line itself: grid 0's image / mask / condition sheets as rank 0 holds them after the
    # per-view shard + exchange must be the bytes a single GPU produces (pure partition: SURVEY §4 item 3)
    def sheet_digest(bufs) -> int:
        acc = 0
        for t in (bufs.image, bufs.mask, bufs.condition):
            acc = (acc * 1000003 + int(t.contiguous().view(torch.int32).to(torch.int64).sum())) & 0xFFFFFFFFFFFFFFFF
        return acc

    step(c2w_d, intr_d)
    torch.cuda.synchronize()
    digest = sheet_digest(sheet.buffers) if rank == 0 else 0
    digest_single = None
    if rank == 0 and N > 1:
        solo = ReferenceSheetRenderer(fld, layout, H, W, ropts, mopts)
        digest_single = sheet_digest(solo(c2w_all[0].to(dev), intr_all[0].to(dev)))
        del solo
    sync_all()

    # e2e: the same unit through the reference-facing plugin API, host buffers in, host result out, every step:
    #   plugin.DatasetGenerator.generate_reference_sheet(graph, 15 reference cameras, w, h)   (datasetgenerator.py:470-593:
    #       15 renders + masks / conditions, paste, Diffuser.diffuse on the sheet, blend, 15 tile cut-outs)
    #   plugin.DatasetGenerator.render_camera(graph, the 16th = dataset camera)                (:677-820)
    # with the cameras in pinned HOST memory (the plugin copies them to the device) and `Diffuser.diffuse` = the one
    # UNet + ControlNet + sampler step of the metric, whose latent is read back to the host.
    import signerf_b200.plugin as P

    class OneStepDiffuser(P.Diffuser):
        """`Diffuser.diffuse(original, rendered, mask, condition)` boundary (diffuser.py:92-106) running ONE step."""
        latent = None

        def diffuse(self, original_image, rendered_image, mask_image=None, condition_image=None):
            if unet is not None:
                OneStepDiffuser.latent = unet.step(original_image, mask_image, condition_image)
            else:
                OneStepDiffuser.latent = original_image
            return original_image

    gcfg = P.DatasetGeneratorConfig(rows=ROWS, cols=COLS, width=W, height=H, downscale_factor=1, fx=float(W), fy=float(W),
                                    cx=W / 2.0, cy=H / 2.0)
    gen = gcfg.setup(original_transform_matrix=torch.eye(4)[:3], original_scale_factor=1.0,
                     transform_poses_to_original_space=lambda x: x, device=dev)
    gen.diffuser = OneStepDiffuser(gcfg.diffuser, dev)
    gen._dist = lambda: (0, 1)                    # every rank runs whole units here (DP over sheets: hot loop #1)
    graph = P.FusedNerfactoGraph(fld, ropts)
    c2w_e2e = c2w_all[rank].contiguous().pin_memory()           # this rank's grid: [16,3,4] in pinned host memory
    res_h = None
    e2e_evs = []
    sync_all()
    for i in range(0 if args.no_e2e else args.warmup + args.steps):
        a, b = ev(), ev()
        a.record()
        cams = P.CameraBatch(c2w_e2e, float(W), float(W), W / 2.0, H / 2.0, W, H)
        gen.generate_reference_sheet(graph, cams[:VIEWS - 1], W, H)
        gen.render_camera(graph, cams[VIEWS - 1])
        out = OneStepDiffuser.latent
        if res_h is None:
            res_h = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        res_h.copy_(out, non_blocking=True)
        b.record()
        b.synchronize()
        if i >= args.warmup:
            e2e_evs.append((a, b))
        flush.zero_()
    sync_all()
    e2e_ms = sum(a.elapsed_time(b) for a, b in e2e_evs) if e2e_evs else float("nan")
    if res_h is None:
        res_h = torch.empty(0)

    t = torch.tensor([total_ms, e2e_ms, k1_ms, unet_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, k1_ms, unet_ms = t.tolist()
    # per-kernel-family split of the diffusion half: one extra EAGER step, every C-ABI call bracketed by CUDA events
    fam = unet.profile_eager() if (unet is not None and rank == 0 and not args.no_e2e) else {}
    loop = unet.full_loop_ms() if (unet is not None and rank == 0 and not args.no_e2e) else None
    codec_ms = None
    if loop is not None:
        # the two ends of the img2img request around the latent loop on this step's sheet:
        # quantise + mask blur + latent mask + VAE encode, and VAE decode + uint8 + overlay compositing
        from signerf_b200 import inpaint as inpaint_mod
        from signerf_b200 import vae as vae_mod
        vcfg = vae_mod.VAEConfig()
        codec = inpaint_mod.A1111InpaintCodec(vae_mod.VAEB200(vcfg, vae_mod.VAERandomWeights(vae_mod.vae_param_schema(vcfg), 2, dev), dev))
        rgb, depth = ops.render_views(fld, c2w_d, intr_d, H, W, ropts)
        mask, cond, _ = ops.mask_condition(c2w_d, intr_d, depth, mopts)
        b_ = sheet.paste(rgb, mask, cond, 0)
        st = codec.prepare(b_.image, b_.mask, b_.condition)
        edited = codec.finish(st.init_latent, st)          # warm-up of both ends
        torch.cuda.synchronize()
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        st = codec.prepare(b_.image, b_.mask, b_.condition)
        e1.record()
        edited = codec.finish(st.init_latent, st)
        e2.record()
        torch.cuda.synchronize()
        if not bool(torch.isfinite(edited).all()):
            raise RuntimeError("non-finite image out of the VAE decoder")
        codec_ms = (e0.elapsed_time(e1), e1.elapsed_time(e2))
        del codec, st, edited

    # C2' (the sampling the reference really uses, and the plugin's default): 256 -> 96 -> 48 cascade through both proposal
    # networks at the benchmark size, timed beside the flat-128 headline (678.6 GB of algorithmic gathers)
    cascade_ms = None
    if rank == 0 and not args.no_e2e:
        fld_c = synthetic.random_field(seed=0, device=dev, dense=True, with_proposals=True)
        copts = ops.RenderOptions(mode="cascade", num_samples=48, num_prop_samples=(256, 96))
        ops.render_views(fld_c, c2w_d, intr_d, H, W, copts)
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        ops.render_views(fld_c, c2w_d, intr_d, H, W, copts)
        b.record()
        torch.cuda.synchronize()
        cascade_ms = a.elapsed_time(b)
        fld_c.close()
        del fld_c

    # BASELINE config 5, generation half (the fine-tune half is §8(f) row 4): the whole product entry point
    # plugin.DatasetGenerator.generate_dataset with masking_mode="shape" (proxy mesh through the CUDA rasteriser),
    # Diffuser(mode="custom") = A1111 pre + VAE encode + 19 UNet+ControlNet evaluations + VAE decode + overlay per sheet,
    # dataset cameras sharded i mod N over the ranks, PNGs + transforms.json written.  Default run: a bounded sample of
    # `--config5-cameras` dataset cameras per rank; `--config5-cameras 30 --config5-total` runs the named 30-camera job.
    config5 = None
    if unet is not None and args.config5_cameras > 0 and not args.no_e2e:
        import shutil
        import tempfile
        from signerf_b200 import inpaint as inpaint_mod
        from signerf_b200 import vae as vae_mod
        n_cam = args.config5_cameras if args.config5_total else args.config5_cameras * N
        vcfg = vae_mod.VAEConfig()
        codec5 = inpaint_mod.A1111InpaintCodec(vae_mod.VAEB200(vcfg, vae_mod.VAERandomWeights(vae_mod.vae_param_schema(vcfg), 2, dev), dev))
        tmp = tempfile.mkdtemp(prefix="sgn_config5_") if rank == 0 else None
        if world > 1:
            box = [tmp]
            dist.broadcast_object_list(box, src=0)
            tmp = box[0]
        g5 = P.DatasetGeneratorConfig(path=tmp, dataset_name="config5", rows=ROWS, cols=COLS, width=W, height=H, downscale_factor=1,
                                      fx=float(W), fy=float(W), cx=W / 2.0, cy=H / 2.0, masking_mode="shape",
                                      diffuser=P.DiffuserConfig(mode="custom"))
        gen5 = g5.setup(original_transform_matrix=torch.eye(4)[:3], original_scale_factor=1.0,
                        transform_poses_to_original_space=lambda x: x, device=dev)
        backend5 = P.InProcessSDXL(unet.net, unet.context, unet.y, codec5)
        gen5.diffuser.attach(backend5)
        # one-time capture of the 19-evaluation trajectory as a CUDA graph (a per-process cost, like loading the weights)
        tcap = time.perf_counter()
        backend5.denoise_latents(unet.init, unet.hint, unet.lat_mask, gen5.diffuser.num_inference_steps,
                                 gen5.diffuser.denoising_strength, gen5.diffuser.guidance_scale,
                                 gen5.diffuser.controlnet_conditioning_scale, gen5.diffuser.seed)
        torch.cuda.synchronize()
        graph_capture_s = time.perf_counter() - tcap
        a5, b5 = ev(), ev()
        a5.record()
        backend5.denoise_latents(unet.init, unet.hint, unet.lat_mask, gen5.diffuser.num_inference_steps,
                                 gen5.diffuser.denoising_strength, gen5.diffuser.guidance_scale,
                                 gen5.diffuser.controlnet_conditioning_scale, gen5.diffuser.seed)
        b5.record()
        torch.cuda.synchronize()
        graphed_loop_ms = a5.elapsed_time(b5)
        pv, pf = synthetic.proxy_mesh(4968)                      # as many triangles as the reference's models/bunny.obj
        gen5.renderer.set_mesh(pv, pf)
        gen5.renderer.scale = [0.01, 0.01, 0.01]                 # x10 "Blender ratio" -> a 0.1-radius object at the origin
        # a transparent random-init field (sigma ~ 0.01): median depth = far plane, so the proxy is in front wherever it
        # covers a pixel and the masks are not empty; rendered with the plugin's default sampling (256 -> 96 -> 48 cascade)
        fld5 = synthetic.random_field(seed=0, device=dev, dense=False, with_proposals=True)
        graph5 = P.FusedNerfactoGraph(fld5)
        ref_c2w, _ = synthetic.camera_ring(VIEWS - 1, W, H)
        syn_c2w, _ = synthetic.camera_ring(n_cam, W, H, phi_deg=(3.0, 357.0))
        # refinement rounds: generation alternates with fine-tuning of the NeRF on the generated images (edit_loop.py)
        from signerf_b200 import edit_loop
        from signerf_b200 import train as train_mod
        emb5 = torch.randn(n_cam + VIEWS, 32, generator=torch.Generator().manual_seed(11))     # one row per generated image
        tuner = edit_loop.FineTuner(train_mod.NerfactoTrainer(fld5, embedding=emb5, pred_normals=synthetic.random_pred_normals()),
                                    num_samples=48,
                                    rays_per_batch=max(1024, 16384 // N), seed=rank)
        rounds, gen_s, train_s, train_losses, frames = [], 0.0, 0.0, [], None
        for rnd in range(args.config5_rounds):
            gen5.dataset_name = f"config5_round{rnd}"
            sync_all()
            t0 = time.perf_counter()
            gen5.generate_dataset(graph5, ref_c2w, synthetic_camera_to_worlds=syn_c2w)
            sync_all()
            t1 = time.perf_counter()
            imgs, cw, it_ = edit_loop.load_generated_images(os.path.join(tmp, gen5.dataset_name))
            losses = tuner.fit(imgs, cw, it_, args.config5_train_steps) if args.config5_train_steps > 0 else []
            sync_all()
            t2 = time.perf_counter()
            gen_s += t1 - t0
            train_s += t2 - t1
            train_losses.append(losses)
            rounds.append({"generate_s": t1 - t0, "fine_tune_s": t2 - t1})
            if rank == 0:
                frames = len(json.load(open(os.path.join(tmp, gen5.dataset_name, "transforms.json")))["frames"])
        t5 = gen_s / max(1, args.config5_rounds)
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)
        per_rank = (n_cam + N - 1) // N
        config5 = {"dataset_cameras": n_cam, "reference_views": VIEWS - 1, "ranks": N, "seconds": t5, "frames_written": frames,
                   "sheets_diffused_per_rank": 1 + per_rank, "seconds_per_sheet": t5 / (1 + per_rank),
                   "dataset_views_per_s": n_cam / t5,
                   "estimated_30_camera_seconds": t5 / (1 + per_rank) * (1 + (30 + N - 1) // N),
                   "denoise_loop_graph_replay_ms": graphed_loop_ms, "denoise_loop_graph_capture_s": graph_capture_s,
                   "refinement_rounds": args.config5_rounds, "fine_tune_steps_per_round": args.config5_train_steps,
                   "rounds": rounds, "total_seconds": gen_s + train_s,
                   "fine_tune_ms_per_step": (train_s / max(1, args.config5_rounds * args.config5_train_steps) * 1e3)
                   if args.config5_train_steps > 0 else None,
                   "fine_tune_losses_rank0": train_losses,
                   "fine_tune_note": "whole nerfacto step (train.NerfactoTrainer): proposal sampler in training mode (256 -> 96 -> 48, "
                                     "stratified + jittered PDF re-sampling), main field + both proposal networks + per-image "
                                     "appearance embeddings + the normal-prediction branch (predict_normals), L1 rgb + "
                                     "interlevel + distortion + orientation + pred-normal losses on 32x32 patches of the "
                                     "generated images, 16 384 rays per step over all ranks, Adam lr 1e-2 on fields and "
                                     "proposal_networks; gradients all-reduced across ranks (one flattened all-reduce); LPIPS (no "
                                     "pretrained weights offline) is not part of the step (DESIGN.md)",
                   "note": "BASELINE config 5, generation half through plugin.DatasetGenerator.generate_dataset: procedural proxy mesh "
                           "with bunny.obj's 4 968 triangles (the reference's mesh file is not shipped), masking_mode='shape', 4x4 sheet of "
                           "512^2 tiles, 20 configured steps at strength 0.9 = 19 UNet+ControlNet evaluations per sheet replayed as ONE CUDA graph, "
                           "PNG encoding + transforms.json included, host wall clock; fine-tune rounds are SURVEY §8(f) row 4"}
        fld5.close()
        del codec5, gen5, graph5, fld5, backend5

    if rank == 0:
        pk = peaks()
        ms_per_step = total_ms / args.steps
        value = N * args.steps / (total_ms / 1e3)
        e2e_value = N * args.steps / (e2e_ms / 1e3)
        k1_bytes = VIEWS * H * W * (SAMPLES * RENDER_BYTES_PER_SAMPLE + RENDER_BYTES_PER_RAY)
        k1_gbs = k1_bytes / (k1_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate (tcgen05 UNet on fp16-representable weights, as the fp16 checkpoints the "
                     "reference up-casts are; f32 hash gather/composite + f16 mma MLP on genuine fp32 NeRF weights in the renderer, "
                     "near-tie median depths re-marched in f32)",
            "data": "synthetic",
            "config": {"workload": "C3 4x4 grid of 512x512 views, flat 128 samples/ray, nerfacto field (16-level 2^19 hash + MLPs), "
                                   "AABB mask + 50x50 dilation + depth condition, 2048x2048 sheet" +
                                   (", + 1 SDXL+ControlNet UNet step (CFG 2)" if unet is not None else "; UNet step NOT YET BUILT (render half only)"),
                       "views": VIEWS, "height": H, "width": W, "samples_per_ray": SAMPLES, "sheet": [layout.height, layout.width],
                       "unet": unet is not None, "sharding": exchange_name,
                       "l2": "256 MiB memset between steps, outside the timed events"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(c2w_e2e.numel() * 4 + VIEWS * 4 * 4),
                    "d2h_bytes_per_step": int(res_h.numel() * res_h.element_size()),
                    "api": "signerf_b200.plugin.DatasetGenerator.generate_reference_sheet(graph, 15 cameras, 512, 512) + "
                           "render_camera(graph, 16th camera); Diffuser.diffuse = one UNet+ControlNet step; cameras from pinned "
                           "host memory, latent read back to the host"},
            "sheet_checksum": {"grid0_u64": digest, "single_gpu_u64": digest_single,
                               "bit_identical_to_single_gpu": (digest == digest_single) if digest_single is not None else None,
                               "what": "wrapping int64 sum of the fp32 bit patterns of grid 0's image / mask / condition sheets on "
                                       "rank 0 after the per-view shard + tile exchange; equals the N=1 line's value"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        render_roof = {"kernel": "k_render_mma", "bound": "hbm", "achieved": k1_gbs, "peak": pk["hbm"], "unit": "GB/s",
                       "frac": k1_gbs / pk["hbm"], "traffic": 97.6e6, "peak_source": pk["source"], "ms_per_launch": k1_ms,
                       "algorithmic_bytes_per_launch": k1_bytes,
                       "note": "hash tables (64 MiB) stay L2-resident: ncu dram bytes per launch are 97.6 MB (profiles/), "
                               "so the algorithmic-gather figure can exceed the HBM copy peak",
                       # the ceiling that actually binds (profiles/r2_ncu_render_summary.txt, r2_k1_wavefronts_per_level.txt):
                       # 2.15e9 LDG.64 per launch touch 4.28 distinct 128-B lines each (levels 12-15: 10-16 lines, 66 % of the
                       # work; levels 0-4: 1.1-1.3 lines, 8 %) and keep the L1 LSU data pipe 80 % busy
                       "l1_ceiling": {"bound": "l1 lsu data pipe (wavefronts)", "frac": 0.803, "ldg_per_launch": 2.1496e9,
                                      "tag_requests_per_ldg": 4.28, "source": "ncu --set full, one launch (profiles/r2_ncu_render_summary.txt)"}}
        if unet is not None:
            tf = UNET_TFLOP_PER_STEP / (unet_ms / 1e3)
            tc = {k: v for k, v in fam.items() if v["flops"] > 0}
            # dominant tensor-bound kernel family by time in the step; k_gemm_tc (linear) and k_attention_tc are within a
            # per cent of each other, so a near-tie (5 %) goes to the one further from its roofline - the line does not flip
            # between runs and reports the weaker number
            top = None
            if tc:
                worst = max(v["ms"] for v in tc.values())
                near = [k for k, v in tc.items() if v["ms"] >= 0.95 * worst]
                top = min(near, key=lambda k: tc[k]["flops"] / tc[k]["ms"])
            line["roofline"] = {
                "kernel": top, "bound": "tensor",
                "achieved": (tc[top]["flops"] / (tc[top]["ms"] / 1e3) / 1e12) if top else None,
                "peak": pk["tensor"], "unit": "TFLOP/s",
                "frac": (tc[top]["flops"] / (tc[top]["ms"] / 1e3) / 1e12 / pk["tensor"]) if top else None,
                # ncu --set full of ONE launch of the dominant shape (2 x 10 heads x 16 384^2, profiles/r2_ncu_attention_v2_summary.txt):
                # dram read + write = 126.13 + 26.13 MB against 168 MB of q / k / v / o (K and V of a head stay in L2)
                "traffic": 152.3e6 if top == "k_attention_tc" else None,
                "peak_source": pk["source"] + " (sustained bf16 cuBLAS)",
                "ms_per_step_in_kernel": tc[top]["ms"] if top else None, "launches_per_step": tc[top]["calls"] if top else None,
                "algorithmic_flops_per_step": tc[top]["flops"] if top else None,
                "unet_step": {"ms": unet_ms, "algorithmic_tflop": UNET_TFLOP_PER_STEP, "achieved": tf,
                              "frac": tf / pk["tensor"], "timed": "CUDA-graph replay inside the timed region"},
                "by_kernel": {k: {"ms": round(v["ms"], 3), "calls": v["calls"],
                                  "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) if v["flops"] else None}
                              for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
                "render": render_roof}
        else:
            line["roofline"] = render_roof
        if loop is not None:
            line["config3_full_inpaint_loop"] = {
                "unet_evaluations": loop[0], "ms": loop[1], "ms_per_evaluation": loop[1] / loop[0],
                "pre_and_vae_encode_ms": codec_ms[0], "vae_decode_and_post_ms": codec_ms[1],
                "render_ms": k1_ms, "total_ms": k1_ms + codec_ms[0] + loop[1] + codec_ms[1],
                "note": "one whole reference-sheet edit (BASELINE config 3): render + A1111 pre (uint8 quantise, mask blur, "
                        "latent mask) + SDXL VAE encode + 20 configured steps at denoising strength 0.9 = 19 UNet+ControlNet CFG "
                        "evaluations on the 2048^2 sheet latent (eager launches) + VAE decode + overlay compositing; "
                        "random-init weights; prompt encoders excluded (once per prompt)"}
        if cascade_ms is not None:
            cb = VIEWS * H * W * (48 * RENDER_BYTES_PER_SAMPLE + 352 * 320 + RENDER_BYTES_PER_RAY)
            line["render_cascade_c2prime"] = {
                "ms": cascade_ms, "algorithmic_bytes": cb, "achieved_gbs": cb / (cascade_ms / 1e3) / 1e9,
                "frac_of_hbm_peak": cb / (cascade_ms / 1e3) / 1e9 / pk["hbm"],
                "note": "16 views 512^2, nerfacto's own sampling (256 -> 96 -> 48 through two proposal networks, then the main "
                        "field on 48 samples): k_prop_weights x2, k_pdf_resample x2, k_render_mma per 4-view chunk; the "
                        "plugin's default mode (FusedNerfactoGraph), not part of the headline unit (flat 128)"}
        if config5 is not None:
            line["config5_generation"] = config5
        if N == 1 and not args.no_cpu_baseline:
            # bounded: a quarter view (65 536 rays, ~15 s) and one 512^2-sheet step (~3 s); `--impl reference` measures the
            # full view and the 1024^2 sheet
            val, cores, sample, parts = cpu_baseline(rays=H * W // 4, render_runs=1, unet_sheet=512, unet_runs=1, with_unet=unet is not None)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "measured_vs_scaled": parts}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
