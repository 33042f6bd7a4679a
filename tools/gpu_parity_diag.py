"""Where does the CUDA path differ from the references?  (GPU box)  Per BSDF / emitter probe: rows and columns outside
rtol 2e-4 and the worst relative error; per scene: fraction of samples whose radiance differs from the oracle's by more
than 1e-3 / 1e-2 / 1e-1 relative.   usage: gpu_parity_diag.py [scene ...]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import abi, nscene
from nori_ray_tracer_b200.gpu import NoriGpu
from oracle_binding import Oracle
G = os.path.join(ROOT, "tests", "golden")
names = sys.argv[1:] or sorted(json.load(open(os.path.join(G, "meta.json")))["scenes"])
g = NoriGpu(0)
BC = ["ev.r", "ev.g", "ev.b", "pdf", "w.r", "w.g", "w.b", "wo.x", "wo.y", "wo.z", "meas", "pdf2"]
for name in names:
    sc = nscene.load_scene(os.path.join(G, f"{name}.nscene"))
    g.upload_scene(sc)
    for b in range(sc.pod.n_bsdfs):
        ref = sc.entries[f"probe.bsdf.{b}.out"]; got = g.probe_bsdf(b, sc.entries[f"probe.bsdf.{b}.in"])
        ok = np.isclose(got, ref, rtol=2e-4, atol=2e-6, equal_nan=True)
        rel = np.abs(got - ref) / (np.abs(ref) + 1e-2)
        rel = np.where(np.isfinite(rel), rel, 0)
        if not ok.all():
            print(f"{name} bsdf {b} type {sc.bsdfs[b].type}: rows ok {ok.all(1).mean():.4f} worst rel {rel.max():.2e} 99.5% {np.quantile(rel.max(1), 0.995):.2e} bad cols",
                  {BC[c]: int((~ok[:, c]).sum()) for c in range(12) if (~ok[:, c]).any()})
    for e in range(sc.pod.n_emitters):
        ref = sc.entries[f"probe.emitter.{e}.out"]; got = g.probe_emitter(e, sc.entries[f"probe.emitter.{e}.in"])
        ok = np.isclose(got, ref, rtol=2e-4, atol=2e-6, equal_nan=True)
        rel = np.abs(got - ref) / (np.abs(ref) + 1e-2); rel = np.where(np.isfinite(rel), rel, 0)
        if not ok.all():
            print(f"{name} emitter {e}: rows ok {ok.all(1).mean():.4f} worst rel {rel.max():.2e} bad cols", {c: int((~ok[:, c]).sum()) for c in range(15) if (~ok[:, c]).any()})
    g.set_option("pool", 1 << 16)
    got = g.render_samples(0, 3, seed=11); want = Oracle(sc, abi).render_samples(0, 3, seed=11)
    rel = (np.abs(got - want) / (np.abs(want) + 1e-3)).max(-1)
    print(f"{name}: samples off by >1e-3 {float((rel > 1e-3).mean()):.2e}  >1e-2 {float((rel > 1e-2).mean()):.2e}  >1e-1 {float((rel > 1e-1).mean()):.2e}  "
          f"mean {got[..., :3].mean():.6f} vs {want[..., :3].mean():.6f}", flush=True)
