#!/usr/bin/env python
"""Turn the golden images the REFERENCE ships (scenes/pa1/ref, pa3/*/ref, pa4/table/ref) into committed fixtures.

    python tests/golden/make_ref_goldens.py            (build container only: reads /root/reference)

For each golden EXR the reference holds for a hot-path integrator this script stores the image box-downsampled by
16 x 16 -- the resolution the image tolerance of SURVEY 8(d) is defined on -- in tests/golden/ref_goldens.npz, and
records in ref_goldens.json which fixture scene (+ integrator override), resolution and sample count reproduce it.
Scenes that have no .nscene fixture yet are exported with oracle/_ref/nori_export (no ray batches / probes).
The photon-mapper goldens (table_pmap, cbox_pmap) are outside the hot path; cbox refs, ajax and
sponza are among the reference's missing large blobs.
"""
import json
import os
import subprocess

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2  # noqa: E402
import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NORI_REFERENCE", "/root/reference")
EXPORT = os.path.join(ROOT, "oracle", "_ref", "nori_export")

# golden (under scenes/), scene xml (under scenes/), fixture .nscene, integrator, spp  (resolution: the golden's own)
GOLDENS = [
    ("pa1/ref/sphere-analytic.exr", "pa1/sphere-analytic.xml", "sphere_analytic_normals", "normals", 32),
    ("pa1/ref/sphere-mesh.exr", "pa1/sphere-mesh.xml", "sphere_mesh_normals", "normals", 32),
    ("pa1/ref/sphere-texture.exr", "pa1/sphere-texture.xml", "sphere_texture_direct", "direct", 32),
    ("pa1/ref/mesh-texture.exr", "pa1/mesh-texture.xml", "mesh_texture_direct", "direct", 32),
    ("pa3/odyssey/ref/odyssey_ems_64spp.exr", "pa3/odyssey/odyssey_ems.xml", "odyssey_mis", "direct_ems", 64),
    ("pa3/odyssey/ref/odyssey_mats_64spp.exr", "pa3/odyssey/odyssey_mats.xml", "odyssey_mis", "direct_mats", 64),
    ("pa3/odyssey/ref/odyssey_mis_32spp.exr", "pa3/odyssey/odyssey_mis.xml", "odyssey_mis", "direct_mis", 32),
    ("pa3/sphere/ref/point_ems.exr", "pa3/sphere/point_ems.xml", "point_ems", "direct_ems", None),
    ("pa3/sphere/ref/sphere_ems.exr", "pa3/sphere/sphere_ems.xml", "sphere_ems", "direct_ems", None),
    ("pa3/sphere/ref/sphere_mats.exr", "pa3/sphere/sphere_mats.xml", "sphere_ems", "direct_mats", None),
    ("pa3/sphere/ref/sphere_mesh_ems.exr", "pa3/sphere/sphere_mesh_ems.xml", "sphere_mesh_ems", "direct_ems", None),
    ("pa3/sphere/ref/sphere2_ems.exr", "pa3/sphere/sphere2_ems.xml", "sphere2_mats", "direct_ems", None),
    ("pa3/sphere/ref/sphere2_mats.exr", "pa3/sphere/sphere2_mats.xml", "sphere2_mats", "direct_mats", None),
    ("pa3/sphere/ref/sphere2_mesh_ems.exr", "pa3/sphere/sphere2_mesh_ems.xml", "sphere2_mesh_ems", "direct_ems", None),
    # veach_ems.xml lights the scene with ANALYTIC spheres, veach_mats / veach_mis.xml with sphere meshes: its own fixture
    ("pa3/veach_mi/ref/veach_ems_256spp.exr", "pa3/veach_mi/veach_ems.xml", "veach_ems", "direct_ems", 256),
    ("pa3/veach_mi/ref/veach_mats_256spp.exr", "pa3/veach_mi/veach_mats.xml", "veach_mis", "direct_mats", 256),
    ("pa3/veach_mi/ref/veach_mis_128spp.exr", "pa3/veach_mi/veach_mis.xml", "veach_mis", "direct_mis", 128),
    ("pa4/table/ref/table_path_mis_512spp.exr", "pa4/table/table_path_mis.xml", "table_path_mis", "path_mis", 512),
    ("pa4/table/ref/table_path_mats_512spp.exr", "pa4/table/table_path_mats.xml", "table_path_mis", "path_mats", 512),
]


def xml_info(path):
    import re
    t = open(path).read()
    spp = int(re.search(r'name="sampleCount"\s+value="(\d+)"', t).group(1))
    integ = re.search(r'<integrator\s+type="(\w+)"', t).group(1)
    return spp, integ


def downsample(img, f=16):
    h, w = (img.shape[0] // f) * f, (img.shape[1] // f) * f
    return img[:h, :w].reshape(h // f, f, w // f, f, -1).mean((1, 3)).astype(np.float32)


def main():
    arrays, meta = {}, {}
    for exr, xml, fixture, integ, spp in GOLDENS:
        img = cv2.imread(os.path.join(REF, "scenes", exr), cv2.IMREAD_UNCHANGED)
        assert img is not None, exr
        img = np.ascontiguousarray(img[..., 2::-1], np.float32)        # BGR -> RGB
        xspp, xinteg = xml_info(os.path.join(REF, "scenes", xml))
        assert xinteg == integ, (xml, xinteg, integ)
        spp = spp or xspp
        nsc = os.path.join(HERE, f"{fixture}.nscene")
        if not os.path.exists(nsc):                                   # geometry without a fixture yet: export it
            subprocess.check_call([EXPORT, os.path.join(REF, "scenes", xml), nsc], stdout=subprocess.DEVNULL)
        key = os.path.splitext(os.path.basename(exr))[0].replace("-", "_")
        arrays[key] = downsample(img)
        meta[key] = {"golden": "scenes/" + exr, "scene": "scenes/" + xml, "fixture": fixture, "integrator": integ, "spp": spp,
                     "res": [int(img.shape[1]), int(img.shape[0])], "mean": float(img.mean())}
        print(key, meta[key])
    np.savez_compressed(os.path.join(HERE, "ref_goldens.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "ref_goldens.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
