#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on one node.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation

A *step* is one full render of BASELINE config 2: the Cornell box (scenes/pa4/cbox), path_mis,
800x600, 1024 spp per GPU = 491.52 M camera paths per GPU per step.  With N > 1 (torchrun, one
process per GPU) every rank renders its own 1024 sample indices [r*1024, (r+1)*1024) of the same
image (weak scaling: per-GPU work fixed, the result is an N*1024-spp image) and the float
accumulation buffers are summed onto rank 0 by ONE NCCL reduce inside the timed region.

metric = Msamples/s (whole job).  `value` is timed on the device (CUDA events on the stream the
kernels run on + the reduce on torch's stream, max over ranks) with the scene resident in HBM;
`e2e` goes through the public host API with host buffers (scene upload + render + film download)
and is timed on the host clock.  One extra profiled step brackets every kernel launch with CUDA
events for the roofline line, and one extra step with traversal counters on measures the
algorithmic bytes per ray (B_ray = 32 B/node + 48 B/primitive + 48 B ray/hit record).
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENE = os.path.join(GOLDEN, "cbox_path_mis.nscene")
WIDTH, HEIGHT, SPP = 800, 600, 1024
WORKLOAD = "cornell-box(pa4/cbox) path_mis 800x600 @1024spp per GPU"


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/r01_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum of one steady-state launch at the bench's pool size)."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(p):
        t = json.load(open(p)).get(kernel)
        if t:
            return float(t["dram_bytes_per_launch"])
    return None


def large_scene_probe(g, host_scene, peak):
    """BASELINE config 4 (10M-triangle height field, path_mis, 3840x2160): a short supplementary measurement of
    the large-scene kernels -- here the scene (0.7 GB) is far larger than L2, so the HBM fraction is a real one.
    Untimed extra of rank 0; NOT part of `value`."""
    import time as _t
    t0 = _t.perf_counter()
    sc = host_scene.heightfield_scene(n=2237)
    build_s = _t.perf_counter() - t0
    g.upload_scene(sc)
    g.set_option("pool", 1 << 22)
    g.render(0, 2, seed=1)
    g.set_option("stats", 1); g.reset_stats(); g.clear_film(); g.render(0, 2, seed=1); kc = g.kernel_stats(); g.set_option("stats", 0)
    g.set_option("kernel_timing", 1); g.reset_stats(); g.clear_film(); g.render(0, 8, seed=1)
    st, ks = g.stats(), g.kernel_stats()
    g.set_option("kernel_timing", 0)
    out = {"workload": "10M-triangle height field (9,999,392 triangles, reference-identical SAH tree) path_mis 3840x2160 @8spp",
           "msamples_per_s": st.samples / st.render_ms / 1e3, "mrays_per_s": st.rays / st.render_ms / 1e3,
           "ms": st.render_ms, "scene_build_s": build_s, "kernel_ms": {k: v["ms"] for k, v in ks.items() if v["ms"]},
           "traversal": "near-child-first order on the 4-wide node layout (options order=2 auto, wide=1)"}
    for k in ("extend", "shadow"):
        c, t = kc[k], ks[k]
        if c["rays"] and t["ms"]:
            b = 32.0 * c["nodes"] / c["rays"] + 48.0 * c["prims"] / c["rays"] + 48.0
            ach = t["rays"] * b / (t["ms"] * 1e-3) / 1e9
            out[f"k_{k}_sm"] = {"boxes_per_ray": c["nodes"] / c["rays"], "prims_per_ray": c["prims"] / c["rays"], "bytes_per_ray": b,
                                "grays_per_s": t["rays"] / t["ms"] / 1e6, "achieved_gbs": ach, "frac_of_measured_hbm": ach / peak}
    return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        mx = max((int(r[1]) for r in self.rows if r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own renderer (oracle/_ref/nori_ref, unmodified
# sources compiled by oracle/Makefile) on the same scene, all host cores through TBB defaults.
# ------------------------------------------------------------------------------------------------
def run_reference_render(spp, width=WIDTH, height=HEIGHT):
    """Render the Cornell box with the reference binary; returns (seconds of its own render timer, cores, kind)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "nori_ref")
    cores = os.cpu_count()
    if os.path.exists(exe):
        src = os.path.join(GOLDEN, "scenes", "cbox")
        with tempfile.TemporaryDirectory() as tmp:
            shutil.copytree(src, os.path.join(tmp, "cbox"))
            xml_path = os.path.join(tmp, "cbox", "cbox_path_mis.xml")
            xml = open(xml_path).read()
            xml = re.sub(r'(name="sampleCount"\s+value=")\d+', rf"\g<1>{spp}", xml)
            xml = re.sub(r'(name="width"\s+value=")\d+', rf"\g<1>{width}", xml)
            xml = re.sub(r'(name="height"\s+value=")\d+', rf"\g<1>{height}", xml)
            open(xml_path, "w").write(xml)
            out = subprocess.run([exe, xml_path], cwd=os.path.join(tmp, "cbox"), capture_output=True, text=True, timeout=3600).stdout
        m = re.search(r"done\. \(took ([0-9.]+)(ms|s|m|h)\)", out)          # the reference's own timer (render.cpp:252)
        if not m:
            raise RuntimeError("reference render failed:\n" + out[-2000:])
        sec = float(m.group(1)) * {"ms": 1e-3, "s": 1.0, "m": 60.0, "h": 3600.0}[m.group(2)]
        return sec, cores, "reference"
    # the reference was not compiled here: time the oracle port instead (all host threads)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from __graft_entry__ import import_package
    import_package()
    from nori_ray_tracer_b200 import abi, nscene
    from oracle_binding import Oracle
    sc = nscene.load_scene(SCENE)
    sc.set_resolution(width, height)
    o = Oracle(sc, abi)
    t = time.perf_counter()
    o.render(0, spp, mode=1)
    return time.perf_counter() - t, cores, "port"


def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    spp = args.ref_spp
    for _ in range(args.warmup):
        run_reference_render(max(1, spp // 4))
    t_total, kind, cores = 0.0, "reference", os.cpu_count()
    for _ in range(args.steps):
        sec, cores, kind = run_reference_render(spp)
        t_total += sec
    samples = WIDTH * HEIGHT * spp
    value = samples * args.steps / t_total / 1e6
    line = {"impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{spp} of 1024 spp per step (per-pass cost is constant)"},
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind,
                             "sample": f"cornell box 800x600 path_mis, {spp} spp per step, TBB on all host cores"},
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp", type=int, default=SPP, help="samples per pixel per GPU per step (headline: 1024)")
    ap.add_argument("--pool", type=int, default=1 << 22)
    ap.add_argument("--ref-spp", type=int, default=16, help="spp of one bounded reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large-scene", action="store_true", help="skip the supplementary 10M-triangle measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from __graft_entry__ import import_package
    import_package()
    from nori_ray_tracer_b200 import host_scene, nscene, render
    from nori_ray_tracer_b200.gpu import NoriGpu

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    sc = nscene.load_scene(SCENE)
    sc.set_resolution(WIDTH, HEIGHT)
    g = NoriGpu(local)
    g.upload_scene(sc)
    g.set_option("pool", args.pool)
    film_t = torch.as_tensor(g.film_device_array(), device=f"cuda:{local}")
    spp = args.spp
    begin = rank * spp                                      # disjoint sample-index (pcg32 initstate) ranges
    samples_per_gpu = WIDTH * HEIGHT * spp
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        g.synchronize()

    def step():
        """render (device-timed by the library's events on its own stream) + the single reduce"""
        g.clear_film()
        g.render(begin, spp, seed=0)
        ms = g.stats().render_ms
        if world > 1:
            ev0.record()
            dist.reduce(film_t, dst=0, op=dist.ReduceOp.SUM)
            ev1.record()
            torch.cuda.synchronize()
            ms += ev0.elapsed_time(ev1)
        return ms

    for _ in range(max(args.warmup, 3)):
        step()
    clocks = ClockSampler(local)
    clocks.start()
    g.reset_stats()
    barrier()
    wall0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        dev_ms += step()
        g.set_option("flush_l2", 256)                        # evict L2 between timed iterations (untimed)
    barrier()
    wall = time.perf_counter() - wall0
    clk = clocks.stop()
    st = g.stats()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * samples_per_gpu / (ms_per_step * 1e-3) / 1e6
    rays_per_sample = st.rays / max(st.samples, 1)

    # ---- end to end through the public host API with host buffers (scene upload, render, film download)
    host_film = np.empty(sc.film_shape, np.float32)
    scene_bytes = sum(int(a.nbytes) for k, a in sc.entries.items() if not k.startswith(("rays", "seq", "probe")))
    barrier()
    e0 = time.perf_counter()
    parts = {"upload": 0.0, "render": 0.0, "reduce": 0.0, "download": 0.0}
    for _ in range(args.steps):
        t0 = time.perf_counter()
        g.upload_scene(sc)                                   # H2D of every scene array (also clears the film)
        t1 = time.perf_counter()
        g.render(begin, spp, seed=0)
        t2 = time.perf_counter()
        if world > 1:
            film_t = torch.as_tensor(g.film_device_array(), device=f"cuda:{local}")
            dist.reduce(film_t, dst=0, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
        t3 = time.perf_counter()
        g.download_film(host_film)                           # D2H of the (H+2b)x(W+2b)x4 accumulation buffer
        t4 = time.perf_counter()
        for k, v in zip(parts, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            parts[k] += 1e3 * v / args.steps
    barrier()
    e2e_s = (time.perf_counter() - e0) / args.steps
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * samples_per_gpu / float(e2e_t.item()) / 1e6

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel: one profiled step (events around every launch) ...
        g.set_option("kernel_timing", 1)
        g.reset_stats(); g.clear_film(); g.render(begin, spp, seed=0)
        ks = g.kernel_stats()
        g.set_option("kernel_timing", 0)
        # ... and one step with the reference's traversal counters on (nodes / primitive tests per ray)
        g.set_option("stats", 1)
        g.reset_stats(); g.clear_film(); g.render(begin, min(spp, 64), seed=0)
        kc = g.kernel_stats()
        g.set_option("stats", 0)
        kernel_ms = {k: v["ms"] for k, v in ks.items()}
        # dominant trace kernel: k_extend (closest-hit + raygen) or k_shade (shading + the any-hit NEE rays)
        trace_dom = max(("extend", "shade"), key=lambda k: kernel_ms[k])
        c = kc[trace_dom]
        b_ray = 32.0 * c["nodes"] / max(c["rays"], 1) + 48.0 * c["prims"] / max(c["rays"], 1) + 48.0
        k = ks[trace_dom]
        peak, peak_src = measured_peak()
        achieved = k["rays"] * b_ray / max(k["ms"] * 1e-3, 1e-12) / 1e9
        roofline = {"bound": "hbm", "kernel": f"k_{trace_dom}", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": measured_traffic(f"k_{trace_dom}"), "peak_source": peak_src,
                    "bytes_per_ray": b_ray, "rays_per_launch": k["rays"] / max(k["launches"], 1),
                    "avg_launch_ms": k["ms"] / max(k["launches"], 1), "kernel_ms_per_step": kernel_ms,
                    "note": "scene (14 primitives, <2 KB) is L1/L2-resident: HBM fraction is small by construction; "
                            "see profiles/ for issue-slot and L1/L2 numbers"}
        cpu = None
        if not args.no_cpu_baseline:
            sec, cores, kind = run_reference_render(args.ref_spp)
            cpu = {"value": WIDTH * HEIGHT * args.ref_spp / sec / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind,
                   "sample": f"cornell box 800x600 path_mis, {args.ref_spp} of 1024 spp, reference binary with TBB on all host cores, "
                             f"its own render timer ({sec:.2f} s)"}
        line = {"metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "spp_per_gpu": spp, "total_spp": spp * world, "pool_slots": args.pool,
                           "l2": "working set (path pool + 7.9 GB sample buffer) >> L2 and L2 flushed between timed steps"},
                "mrays_per_s": value * rays_per_sample, "rays_per_sample": rays_per_sample,
                "wall_ms_per_step": 1e3 * wall / args.steps,
                "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": scene_bytes,
                        "d2h_bytes_per_step": int(host_film.nbytes), "ms_per_step": 1e3 * e2e_s,
                        "host_ms_breakdown": {k: round(v, 3) for k, v in parts.items()}},
                "gpu_launches": int(st.kernel_launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu}
        if world == 1 and not args.no_large_scene:
            try:
                line["large_scene"] = large_scene_probe(g, host_scene, peak)
            except Exception as e:                                   # supplementary: never fail the headline line
                line["large_scene"] = {"error": str(e)[:200]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
