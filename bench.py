#!/usr/bin/env python
"""bench.py — particle-updates/sec of one full CUBEP3M `particle_mesh` step (PM + PP) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c0|tiny]

A "step" is one pass of the hot path (drift -> cell sort -> particle pass -> per-tile fine mesh -> PP -> coarse mesh
-> ghost deletion) over one synthetic LCDM box.  Workload at N=1 is BASELINE.json configs[2], the largest single-GPU configuration:
512^3 particles on a 1024^3 fine mesh, PPINT + PP_EXT, nodes_dim=1, tiles_node_dim=4 (nf_tile=304), one B200; at N>1 every GPU owns one
such node (configs[4]: weak scaling at 512^3 particles per GPU; N=8 is configs[3]'s 1024^3 particles on a 2048^3 mesh, nodes_dim=2).
`--workload c1` is configs[1] (256^3 particles, PPINT only), `c0` configs[0].
`value`  : particles / device-seconds per step with the particles resident in HBM (CUDA events on the library's stream).
`e2e`    : the same step through the C ABI in strict drop-in mode: pinned-host xv -> H2D, particle_mesh, D2H.
`roofline`: dominant kernel class, algorithmic bytes per launch / its mean device time (events around every launch).
`cpu_baseline` / `--impl reference`: the CPU oracle (C++/OpenMP restatement of the reference; the Fortran itself cannot
be built in this image) timed on this box's host cores, all of them (OMP threads set explicitly: torchrun exports OMP_NUM_THREADS=1).
Its FFT is a vectorised batched Stockham transform (oracle/fft_ref.h), not FFTW: a stated baseline, not the target.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cubep3m_b200 import default_config, ic  # noqa: E402
from cubep3m_b200.abi import max_np  # noqa: E402
from cubep3m_b200 import topology as topo  # noqa: E402

WORKLOADS = {
    # name: (nf_tile, tiles_node_dim, ppint, pp_ext, box Mpc/h, z_i, description)
    "c1": (304, 2, 1, 0, 200.0, 100.0, "BASELINE configs[1]: 256^3 particles, 512^3 fine mesh, PPINT on, nodes_dim=1, tiles_node_dim=2 (nf_tile=304)"),
    "c1a": (176, 4, 1, 0, 200.0, 100.0, "BASELINE configs[1] variant: 256^3 particles, 512^3 fine mesh, PPINT on, tiles_node_dim=4 (nf_tile=176)"),
    "c2": (304, 4, 1, 1, 200.0, 100.0, "BASELINE configs[2]: 512^3 particles, 1024^3 fine mesh, PPINT + PP_EXT, nodes_dim=1, tiles_node_dim=4 (nf_tile=304); ICs = 2x2x2 periodic replication of the 256^3-particle box"),
    "c0": (176, 2, 0, 0, 200.0, 100.0, "BASELINE configs[0]: 128^3 particles, 256^3 fine mesh, PM only, tiles_node_dim=2 (nf_tile=176)"),
    "c1c": (560, 1, 1, 0, 200.0, 100.0, "BASELINE configs[1] variant: 256^3 particles, 512^3 fine mesh, PPINT on, tiles_node_dim=1 (nf_tile=560)"),
    "c1x": (304, 2, 1, 1, 200.0, 100.0, "BASELINE configs[1] box (256^3 particles, 512^3 fine mesh, nf_tile=304) with PPINT + PP_EXT on: one octant of configs[2], used for the clustered-input PP measurement"),
    "c0x": (176, 2, 1, 1, 200.0, 100.0, "profiling aid: BASELINE configs[0] box (128^3 particles, 256^3 fine mesh) with PPINT + PP_EXT on"),
    "tiny": (112, 2, 1, 0, 50.0, 20.0, "dev smoke: 64^3 particles, 128^3 fine mesh"),
}


def make_cfg(name):
    n, T, ppint, pp_ext, box, z_i, desc = WORKLOADS[name]
    cfg = default_config(nf_tile=n, tiles_node_dim=T, ppint=ppint, pp_ext=pp_ext)
    return cfg, box, z_i, desc


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(cfg, np_local, np_all):
    """Algorithmic bytes PER LAUNCH of each kernel class.  A = one padded real tile."""
    n, m, fd = cfg.nf_tile, cfg.m, cfg.m + 3
    A = 4.0 * (n + 2) * n * n
    r = fd / n
    NF = cfg.H ** 3 * 64
    N = cfg.nc_dim
    ghosts = max(np_all - np_local, 0)
    return {
        # per-particle figures are SURVEY §8(d)'s: drift 36 B; key + sort 64 B (12 read, 4 key | 24 + 24 permute); pass ~48 B per ghost;
        # coarse deposit 12 B; compaction 48 B per surviving particle (the coarse kick is fused into it, its 8-point gather hits L2)
        "drift": 36.0 * np_local,                       # 24 B read + 12 B written per particle
        "key_hist": (12.0 + 4.0) * np_all,              # position read, key written (the cell-table atomics are this design's cost, not the algorithm's)
        "scan": 2.0 * NF,                               # design-dependent, per launch: 4 launches per step move 8 B per fine cell of the table (2-byte counts read twice, 4-byte starts written)
        "scatter": (4.0 + 24.0 + 24.0) * np_all,        # key, record read, record written
        "pass_pack": 36.0 * np_local / 3 + 24.0 * ghosts / 3,   # per launch (3 axes): the fused drift of the first axis + the packed ghosts, averaged
        "pass_unpack": 48.0 * ghosts / 3,
        "ngp_density": None,                             # tile_counts + delta-list kernels (tiny)
        "fft_x_r2c": 8.0 * (n - 8) ** 3 + A,             # fused NGP deposit: 2 table entries per deposited cell read, half-spectra written
        "fft_fwd_strided": 2.0 * A,                      # y pass in place
        "fft_inv_z_mul": A + 1.5 * A + 3 * r * A,        # fused z pass: spectrum + 3 kernel components read, 3 cropped-z results written
        "fft_inv_y": 3 * (r * A + r * r * A),            # 3 components, cropped planes read, cropped rows written
        "fft_x_c2r": 3 * (r * r * A + 4.0 * fd ** 3),    # 3 components, cropped rows read, force cube written (+ max |F|^2 fused)
        "force_max": None,
        "ngp_kick": 48.0 * np_local / cfg.tiles_node,   # §8(d): 12 pos + 12 force gather + 12 + 12 velocity RMW per particle
        "cic_mass": 12.0 * np_all * ((cfg.nc_node + 2) / cfg.H) ** 3 + 4.0 * cfg.nc_node ** 3,   # §8(d): 12 B per deposited particle (+ rho_c written once)
        "cic_kick": 48.0 * np_local,                    # fused coarse kick + compaction: 24 B record read + 24 B written
        "compact": None,
        "coarse_fft": None, "coarse_misc": None, "coarse_xchg": None, "ppint": None, "ppext": None, "ppext_dense": None, "ppext_margin": None, "misc": None,
        "_NF": NF, "_A": A,
    }


def fp32_peak():
    """FFMA peak measured with tools/fp32_peak.cu on this pool's B200 (profiles/r1_fp32_peak.json); nominal 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4."""
    p = os.path.join(ROOT, "profiles", "r1_fp32_peak.json")
    try:
        d = json.load(open(p))
        for k in ("ffma_tflops", "FFMA", "ffma"):
            if k in d:
                return float(d[k]), "measured (profiles/r1_fp32_peak.json)"
    except Exception:
        pass
    return 67.7, "measured in round 1 (tools/fp32_peak.cu, FFMA)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def make_ics(cfg, box, z_i, seed=12345):
    """Zel'dovich ICs for one node; meshes beyond 512^3 are built by periodic replication of a 512^3-mesh box (host FFT cost)."""
    t = time.time()
    nc = cfg.mT
    base = min(nc, 512)
    xv = ic.zeldovich_ics(base, box=box, z_i=z_i, seed=seed)
    if nc > base:
        xv = ic.tile_box(xv, base, nc // base)
    return xv, time.time() - t


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function) BEFORE the pinned host buffers are allocated,
    so that the e2e leg's H2D / D2H copies of every rank stay on the GPU's own NUMA node (round 1: eight ranks staged through one node and the
    e2e weak-scaling efficiency at 8 GPUs was 0.33). Returns a note for the JSON line; a no-op when sysfs has nothing to say."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        path = f"/sys/bus/pci/devices/{bdf}/local_cpulist"
        txt = open(path).read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return f"rank bound to the {len(allowed)} CPUs local to GPU {bdf}"
        return f"GPU {bdf}: all allowed CPUs are local ({txt})"
    except Exception as e:      # noqa: BLE001
        return f"no NUMA binding ({type(e).__name__})"


def oracle_run(cfg, xv, z_i, steps, warmup, budget_s=None):
    """Times the CPU oracle (all host threads) on full particle_mesh steps of the given workload; with budget_s the number of timed steps
    is cut so that the run stays inside the budget (at least one)."""
    from oracle import Oracle
    from cubep3m_b200.lib import clock_init, timestep, absorb_limiters
    o = Oracle(cfg, threads=host_threads())
    t_begin = time.perf_counter()
    o.set_particles(xv)
    clk = clock_init(z_i, ppint=cfg.ppint, pp_ext=cfg.pp_ext)
    rng = np.random.default_rng(777)
    shake = np.zeros(3, np.float32)
    mass_p = float(np.float32(cfg.nf_physical_dim) ** 3 / np.float32(len(xv)))
    times, stages = [], None
    for s in range(warmup + steps):
        timestep(clk)
        off = ((rng.random(3, dtype=np.float32) - np.float32(0.5)) * np.float32(16.0) - shake).astype(np.float32)
        shake = shake + off
        t = time.perf_counter()
        out = o.particle_mesh(clk.dt, clk.dt_old, clk.a_mid, mass_p, off)
        dtm = time.perf_counter() - t
        absorb_limiters(clk, out)
        if s >= warmup:
            times.append(dtm)
            stages = out.stages()
        if budget_s is not None and times and (time.perf_counter() - t_begin) + 1.2 * dtm > budget_s:
            break
    threads = o.threads
    o.close()
    return float(np.mean(times)), threads, stages, len(times)


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the Fortran cannot be compiled here) on the FULL workload box, rank 0 only,
    all host threads; bounded by running fewer steps (never a smaller box): one warm-up step if a step is short, then as many of the K
    steps as fit ~150 s."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, box, z_i, desc = make_cfg(args.workload)
    xv, _ = make_ics(cfg, box, z_i)
    steps, warmup = args.steps, args.warmup
    big = len(xv) > 64 * 1024 * 1024
    cap_warm = 0 if big else min(warmup, 1)
    sec, threads, stages, ran = oracle_run(cfg, xv, z_i, max(1, steps), cap_warm, budget_s=150.0)
    val = len(xv) / sec
    line = {
        "metric": "particle_updates_per_sec", "value": val, "unit": "particles/s", "impl": "reference", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "particles_per_gpu": int(len(xv)), "timed_steps_actually_run": ran, "warmup_actually_run": cap_warm,
                   "note": "one node of the workload (what one GPU owns) on this box's host cores; the step count, not the box, is what bounds the run"},
        "cpu_baseline": {"value": val, "unit": "particles/s", "cores": threads, "kind": "port",
                         "sample": f"{ran} full particle_mesh step(s) of the whole workload box after {cap_warm} warm-up, {threads} OpenMP threads "
                                   "(oracle port of the reference; FFT = oracle/fft_ref.h batched Stockham, not FFTW)"},
        "e2e": {"value": val, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "stages_ms": stages,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cubep3m_b200.lib import ParticleMesh, clock_init, timestep, absorb_limiters

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_note = bind_to_gpu_numa_node(local_rank) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg, box, z_i, desc = make_cfg(args.workload)
    cfg.local_gpu = local_rank
    cfg.rank = rank
    grid = topo.grid_for_world(world)
    nccl_id = None
    if world > 1:
        # weak scaling: every GPU owns one cubic node of the N=1 workload; the rank grid is (2,1,1)/(2,2,1)/(2,2,2) (cubep3m_b200/topology.py)
        for i in range(3):
            cfg.nodes_dim_xyz[i] = grid[i]
        from cubep3m_b200.lib import get_unique_id
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(get_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())
    # every rank holds the same periodic box (same seed): the global field is its periodic replication, continuous across ranks
    xv, t_ic = make_ics(cfg, box, z_i, seed=12345)
    npart = len(xv)
    mass_p = float(np.float32(cfg.mT) ** 3 / np.float32(npart))
    pm = ParticleMesh(cfg, nccl_id=nccl_id, world_size=world)
    host = torch.empty((max_np(cfg), 6), dtype=torch.float32).pin_memory().numpy()
    host[:npart] = xv
    pm.upload_particles(host[:npart])
    if world > 1:
        dist.barrier()          # pinning the host buffers takes seconds and varies per rank: start the first step together
    clk = clock_init(z_i, ppint=cfg.ppint, pp_ext=cfg.pp_ext)
    rng = np.random.default_rng(777)
    shake = np.zeros(3, np.float32)

    def one_step(strict=False):
        nonlocal shake, npart
        timestep(clk)
        off = ((rng.random(3, dtype=np.float32) - np.float32(0.5)) * np.float32(16.0) - shake).astype(np.float32)   # update_position.f90:56-58
        shake = shake + off
        if strict:
            pm.upload_particles(host[:npart])
        out = pm.particle_mesh(clk.dt, clk.dt_old, clk.a_mid, mass_p, off)
        if strict:
            got = pm.download_particles(out=host)
            npart = len(got)
        absorb_limiters(clk, out)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    evolve = None
    if args.evolve_to_z is not None:
        # clustered input: the box is evolved on the GPU from the ICs to redshift z with the driver twin choosing dt, so that PPINT /
        # PP_EXT are measured on halos instead of the z_i near-lattice; the timed steps that follow continue the same run
        a_stop = 1.0 / (1.0 + args.evolve_to_z)
        t_ev = time.perf_counter()
        n_ev = 0
        while clk.a < a_stop and n_ev < args.evolve_max_steps:
            o_ev = one_step()
            n_ev += 1
            if rank == 0 and n_ev % 100 == 0:
                print(f"[evolve] step {n_ev} a={clk.a:.4f} z={1 / clk.a - 1:.2f} dt={clk.dt:.4f} step_ms={o_ev.stage_ms[12]:.2f} pp_ext_ms={o_ev.stage_ms[7]:.2f} "
                      f"limiters f={o_ev.dt_f_acc:.3f} pp={o_ev.dt_pp_acc:.3f} ppx={o_ev.dt_pp_ext_acc:.3f} c={o_ev.dt_c_acc:.3f}", file=sys.stderr, flush=True)
        evolve = {"to_z": 1.0 / clk.a - 1.0, "steps": n_ev, "seconds": time.perf_counter() - t_ev}
    for _ in range(args.warmup):
        one_step()
    # ---- timed region A: K steps, resident mode, no per-launch instrumentation -> `value`
    pm.set_profiling(False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = pm.launches
    t0 = time.perf_counter()
    dev_ms, last = 0.0, None
    for _ in range(args.steps):
        last = one_step()
        dev_ms += last.stage_ms[12]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = pm.launches - l0
    # `ms_per_step` is the host clock around the K steps, bracketed by barrier + cudaDeviceSynchronize on both sides (every step ends with
    # a device->host read of the limiters, so there is no queued work outside the bracket); the sum of the library's own CUDA-event step
    # times is reported beside it as device_ms_per_step (it excludes the host-side timestep logic between steps).
    dev_step = dev_ms / args.steps
    ms_step = wall_ms / args.steps
    np_all = last.np_with_ghosts
    # ---- timed region B: the same K steps continued with every launch bracketed by CUDA events on its stream -> per-kernel-class
    # durations for the roofline (the instrumentation itself costs ~6 % of a step, which is why `value` comes from region A)
    class_ms, prof_ms = {}, 0.0
    pp_i, pp_e = pm.pair_counts()
    pair_counts = {"ppint": pp_i, "ppext": pp_e}
    tile_streams = int(os.environ.get("CUBEP3M_B200_TILE_STREAMS", "2"))
    if not args.no_profile:
        pm.set_tile_streams(1)         # one fine tile in flight: per-kernel event times are only unambiguous without tile overlap
        pm.set_profiling(True)
        barrier()
        for _ in range(args.steps):
            o2 = one_step()
            prof_ms += o2.stage_ms[12]
            for k, (ms, nl) in pm.kernel_times().items():
                a = class_ms.setdefault(k, [0.0, 0])
                a[0] += ms; a[1] += nl
        barrier()
        pm.set_profiling(False)
        pm.set_tile_streams(min(tile_streams, cfg.tiles_node))
        prof_ms /= args.steps
    # the clock sampler covers the timed region and the instrumented one (the same steps, back to back, both under load)
    clocks = sampler.stop() if sampler else None
    # ---- e2e: strict drop-in mode through the C ABI with host buffers
    host[:npart] = pm.download_particles()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        one_step(strict=True)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps

    # ---- the acceptance metric on the benchmark box itself: cic_power of the resident particles on the global nf_physical_dim^3 mesh (1024^3 at N = 1,
    # 2048^3 over 8 GPUs: z-slabs / y-pencils, four-step transforms). Collective over the ranks; needs a cubic rank grid (N = 1 or 8). Not part of `value`.
    power_info = None
    if grid[0] == grid[1] == grid[2]:
        try:
            barrier()
            t0 = time.perf_counter()
            pk, pd2, _ = pm.cic_power(box * grid[0], shake=shake)
            barrier()
            power_info = {"mesh": int(cfg.mT * grid[0]), "ms": (time.perf_counter() - t0) * 1e3, "k_first": float(pk[0]), "delta2_first5": [float(v) for v in pd2[:5]],
                          "delta2_nyquist": float(pd2[-1]), "finite": bool(np.isfinite(pd2).all())}
        except Exception as e:      # noqa: BLE001 - the bench line must survive
            power_info = {"error": f"{type(e).__name__}: {e}"}
    # ---- the halo finder's density + maxima pass (halofind.f90:564-672) over the node's tiles, in the state a halofind step is in (cubepm.f90:193-198:
    # link_list, particle_pass, halofind, delete_particles). One rank only (the bench line must not depend on it); not part of `value`.
    halo_info = None
    if world == 1:
        try:
            pm.link_list(); pm.particle_pass()
            torch.cuda.synchronize()
            halo_info = {"tiles": int(cfg.tiles_node_dim) ** 3, "den_peak_cutoff": 100.0}
            for scheme, ngph in (("ngp", True), ("cic", False)):     # -DNGPH (every maintained build) and the fine_cic_mass branch
                t0 = time.perf_counter()
                pks, cft = pm.halofind_peaks(mass_p, 100.0, True, ngph)
                halo_info[scheme] = {"ms": (time.perf_counter() - t0) * 1e3, "n_peaks": int(len(pks)),
                                     "clumping_factor": float(cft[1] * float(cfg.mT) ** 3 / max(cft[0] ** 2, 1e-300))}
            pm.delete_particles()
        except Exception as e:      # noqa: BLE001 - the bench line must survive
            halo_info = {"error": f"{type(e).__name__}: {e}"}
    if world > 1:
        t = torch.tensor([ms_step, e2e_ms, dev_step], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms, dev_step = [float(v) for v in t.tolist()]
        tot = torch.tensor([float(last.np_local)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot)
        total_particles = float(tot.item())
    else:
        total_particles = float(last.np_local)

    if rank == 0:
        peak, peak_src = read_peaks()
        ab = algorithmic_bytes(cfg, last.np_local, np_all)
        stages = {}
        for k, (ms, nl) in class_ms.items():
            if nl == 0:
                continue
            per = ms / nl
            e = {"ms_per_step": ms / args.steps, "launches_per_step": nl / args.steps, "us_per_launch": per * 1e3,
                 "share_of_step": (ms / args.steps) / max(prof_ms, 1e-9)}
            if ab.get(k):
                e["algorithmic_MB_per_launch"] = ab[k] / 1e6
                e["achieved_GBs"] = ab[k] / (per * 1e-3) / 1e9
                e["frac_of_hbm_peak"] = e["achieved_GBs"] / peak
            if k == "ppext" and "ppext_dense" in class_ms:
                ms = ms + class_ms["ppext_dense"][0]     # the pair count covers the sparse (tiled) and the dense (cell-pair) kernels together
                e["ms_per_step_incl_dense"] = ms / args.steps
            if k in pair_counts and pair_counts[k] > 0:
                # FP32-pipe stages: 20 flop per ordered pair interaction evaluated against the measured FFMA peak
                fpk, fsrc = fp32_peak()
                sec = (ms / args.steps) * 1e-3
                e["ordered_pairs_per_step"] = pair_counts[k]
                e["pairs_per_s"] = pair_counts[k] / sec
                e["achieved_TFLOPs"] = 20.0 * pair_counts[k] / sec / 1e12
                e["frac_of_fp32_peak"] = e["achieved_TFLOPs"] / fpk
                e["fp32_peak_TFLOPs"] = fpk
                e["fp32_peak_source"] = fsrc
            stages[k] = e
        if not stages:
            stages = {"fft_inv_z_mul": {"ms_per_step": 0.0, "launches_per_step": 0, "us_per_launch": 0.0, "share_of_step": 0.0, "achieved_GBs": 0.0}}
        dom = max((k for k in stages if ab.get(k)), key=lambda k: stages[k]["ms_per_step"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(dom)
            except Exception:
                traffic = None
        roofline = {"kernel": dom, "bound": "hbm", "achieved": stages[dom]["achieved_GBs"], "peak": peak, "unit": "GB/s",
                    "frac": stages[dom]["achieved_GBs"] / peak, "traffic": traffic, "peak_source": peak_src,
                    "share_of_step": stages[dom]["share_of_step"], "launches_per_step": stages[dom]["launches_per_step"],
                    "instrumented_ms_per_step": prof_ms,
                    "note": "algorithmic bytes per launch / mean CUDA-event time per launch over K instrumented steps that directly follow the K timed "
                            "steps (instrumented with ONE fine tile in flight so that a launch's event time is its own; the timed steps keep two tiles "
                            "in flight; the coarse-mesh solve overlaps on its own stream in both, so a fine-mesh launch's event time can include SM "
                            "time lent to coarse kernels). With PP_EXT on, the largest single kernel of the step is pp::ppext_tiled_kernel, which is "
                            "FP32-issue-bound, not HBM-bound: its pairs/s and fraction of the measured FFMA peak are in stages.ppext (roofline_fp32 "
                            "repeats them); this object covers the largest HBM-bound kernel class."}
        roofline_fp32 = None
        if "ppext" in stages and stages["ppext"].get("frac_of_fp32_peak") is not None:
            e = stages["ppext"]
            roofline_fp32 = {"kernel": "ppext (pp::ppext_tiled_kernel + dense-cell kernels)", "bound": "fp32", "achieved": e["achieved_TFLOPs"], "peak": e["fp32_peak_TFLOPs"],
                             "unit": "TFLOP/s", "frac": e["frac_of_fp32_peak"], "pairs_per_s": e["pairs_per_s"], "share_of_step": e["share_of_step"],
                             "note": "20 flop per ordered pair (SURVEY 8d) / measured FFMA peak (tools/fp32_peak.cu)"}
        cpu = None
        if world == 1 and not args.no_cpu:
            # bounded sample: for boxes beyond 256^3 particles one octant-sized node of the same workload (same tile size, flags, density and
            # ICs: the 256^3-particle box the workload's ICs replicate), which is 1/8 of the work at the same per-particle cost
            scfg, sxv, snote = cfg, xv, "the whole workload box"
            if npart > 64 * 1024 * 1024:
                scfg = default_config(nf_tile=cfg.nf_tile, tiles_node_dim=cfg.tiles_node_dim // 2, ppint=cfg.ppint, pp_ext=cfg.pp_ext)
                sxv, _ = make_ics(scfg, box, z_i)
                snote = (f"a {round(len(sxv) ** (1 / 3))}^3-particle node of the same workload (1/8 of the box: same nf_tile, flags, density and ICs; its "
                         f"{scfg.tiles_node_dim ** 3} tiles leave part of the host threads idle in the tile loop - the reference arm, --impl reference, times the whole box)")
            sec, threads, ost, ran = oracle_run(scfg, sxv, z_i, 3, 0, budget_s=30.0)
            cpu = {"value": len(sxv) / sec, "unit": "particles/s", "cores": threads, "kind": "port", "ms_per_step": sec * 1e3,
                   "sample": f"{ran} full particle_mesh step(s) from the ICs on {snote}, {threads} OpenMP threads (oracle FFT: batched Stockham, not FFTW)",
                   "stages_ms": {k: round(v, 1) for k, v in ost.items()}}
        line = {
            "metric": "particle_updates_per_sec", "value": total_particles / (ms_step * 1e-3), "unit": "particles/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "particles_per_gpu": int(npart), "particles_with_ghosts": int(np_all),
                       "timing": "host clock around K steps between barrier+synchronize, max over ranks (CUDA-event sum in device_ms_per_step); working set (particles 0.4 GB + cell table 1.4 GB) exceeds the 126 MB L2",
                       "ics": f"Zel'dovich LCDM (EH no-wiggle), z_i={z_i}, box={box} Mpc/h per node, numpy seed 12345 (same box on every rank), generated in {t_ic:.1f}s",
                       "rank_grid": list(grid), "parallelism": f"{world} rank(s), one cubic node of {cfg.tiles_node} tiles per GPU; particle_pass packed straight into the neighbour's memory over NVLink (NCCL send/recv fallback); coarse mesh " + ("on the one GPU" if world == 1 else "slab-decomposed, pencil transposes and cube/halo scatters stored straight into the peers' memory (all-gather + replicated solve as fallback)"),
                       "mode": "resident (particles stay in HBM between steps)", "fine_tiles_in_flight": min(tile_streams, cfg.tiles_node),
                       "evolved": evolve, "ppext_blocks_tiled_fallback": list(pm.ppext_blocks())},
            "device_ms_per_step": dev_step,
            "e2e": {"value": total_particles / (e2e_ms * 1e-3), "unit": "particles/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(npart) * 24, "d2h_bytes_per_step": int(npart) * 24, "steps": e2e_steps,
                    "mode": "strict drop-in: pinned host xv -> H2D, particle_mesh, D2H every step (cubepm.f90:143 semantics)", "host_numa": numa_note},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_fp32": roofline_fp32, "cpu_baseline": cpu,
            "stages": stages,
            "cic_power": power_info,
            "halofind_peaks": halo_info,
            "stage_ms_last_step": {k: round(v, 3) for k, v in last.stages().items()},
            "limiters_last_step": {"dt_f_acc": last.dt_f_acc, "dt_pp_acc": last.dt_pp_acc, "dt_c_acc": last.dt_c_acc,
                                   "sum_rho_f": last.sum_rho_f, "sum_rho_c": last.sum_rho_c, "a": clk.a, "nts": clk.nts},
        }
        print(json.dumps(line), flush=True)
    pm.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE JSON line: everything libraries print while the run is set up (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--evolve-to-z", type=float, default=None, help="evolve the box on the GPU to this redshift before measuring (clustered input for the PP stages)")
    ap.add_argument("--evolve-max-steps", type=int, default=4000)
    ap.add_argument("--no-profile", action="store_true", help="do not bracket launches with events in the timed region (overhead check)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
