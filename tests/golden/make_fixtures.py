#!/usr/bin/env python
"""Regenerate tests/golden/ from the reference (run in the build container, where /root/reference exists).

    python tests/golden/make_fixtures.py [name ...]

For every fixture below this script
  1. copies the scene directory out of /root/reference/scenes (read-only) into a scratch dir and
     rewrites resolution / sample count / integrator in the XML (nothing else),
  2. runs oracle/_ref/nori_export  (reference parser + OBJ loader + SAH BVH builder, unmodified) to
     write <name>.nscene: the flat scene description, plus ray batches answered by the reference's
     own BVH::rayIntersect and the reference's per-sample sequence for block (0,0),
  3. runs oracle/_ref/nori_ref (the reference's headless front end, unmodified) and stores the
     rendered image as <name>.ref<spp>.npy (float32 H x W x 3),
  4. for <test type="ttest"> files: splits the test into its <scene> children and records the
     known answers from the XML in meta.json.
The GPU box has no /root/reference: tests there only read what this script wrote.
"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import xml.etree.ElementTree as ET

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2  # noqa: E402
import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NORI_REFERENCE", "/root/reference")
EXPORT = os.path.join(ROOT, "oracle", "_ref", "nori_export")
NORI = os.path.join(ROOT, "oracle", "_ref", "nori_ref")

# name, source xml (under scenes/), overrides, export options, reference renders (spp list)
FIXTURES = [
    dict(name="cbox_path_mis", src="pa4/cbox/cbox_path_mis.xml", res=(200, 150), rays=12000, seq=2000, ref_spp=[4, 128], variance=True),
    dict(name="cbox_path_mats", src="pa4/cbox/cbox_path_mats.xml", res=(200, 150), rays=0, seq=1000, ref_spp=[4, 256]),
    dict(name="sphere_mesh_normals", src="pa1/sphere-mesh.xml", res=(128, 128), rays=8000, seq=1000, ref_spp=[4], variance=True),
    dict(name="sphere_analytic_normals", src="pa1/sphere-analytic.xml", res=(128, 128), rays=4000, seq=1000, ref_spp=[4]),
    dict(name="sphere_ems", src="pa3/sphere/sphere_ems.xml", res=(128, 128), rays=0, seq=1000, ref_spp=[4]),
    dict(name="sphere2_mats", src="pa3/sphere/sphere2_mats.xml", res=(128, 128), rays=0, seq=1000, ref_spp=[4, 256]),
    dict(name="point_ems", src="pa3/sphere/point_ems.xml", res=(128, 128), rays=0, seq=1000, ref_spp=[4]),
    dict(name="veach_mis", src="pa3/veach_mi/veach_mis.xml", res=(192, 128), rays=8000, seq=1000, ref_spp=[4, 64]),
    dict(name="odyssey_mis", src="pa3/odyssey/odyssey_mis.xml", res=(192, 108), rays=0, seq=1000, ref_spp=[4]),
    dict(name="table_path_mis", src="pa4/table/table_path_mis.xml", res=(200, 150), rays=12000, seq=1000, ref_spp=[4, 64]),
    dict(name="disney_cbox", src="project/disney/cbox_path_mis.xml", res=(200, 150), rays=0, seq=1000, ref_spp=[4, 64]),
    dict(name="volumetric", src="project/volumetric/volumetric.xml", res=(200, 150), rays=0, seq=1000, ref_spp=[4, 64]),
    dict(name="spotlight_direct", src="project/spotlight/sphere-texture.xml", res=(128, 128), rays=0, seq=1000, ref_spp=[4]),
    dict(name="sphere_texture_direct", src="pa1/sphere-texture.xml", res=(128, 128), rays=0, seq=1000, ref_spp=[4]),
    dict(name="sphere_av", src="pa1/sphere-mesh.xml", res=(96, 96), rays=0, seq=1000, ref_spp=[4],
         integrator=("av", '<float name="length" value="0.5"/>')),
    # authored variants (same geometry, other hot-path plugins): thin-lens camera, spot light + path_mis
    dict(name="cbox_thinlens", src="pa4/cbox/cbox_path_mis.xml", res=(200, 150), rays=0, seq=1000, ref_spp=[4, 64],
         camera=("thinlens", '<float name="lensRadius" value="0.05"/><float name="focalDist" value="4.6"/>')),
    dict(name="cbox_spot_point_mis", src="pa4/cbox/cbox_path_mis.xml", res=(200, 150), rays=0, seq=1000, ref_spp=[4, 64],
         extra='<emitter type="spotlight"><point name="position" value="0,1.5,0.5"/><color name="color" value="30,20,10"/>'
               '<vector name="direction" value="0.2,-1,-0.1"/><float name="falloffStart" value="15"/><float name="totalWidth" value="35"/></emitter>'
               '<emitter type="point"><point name="position" value="-0.5,0.8,0.6"/><color name="power" value="4,8,12"/></emitter>'),
    dict(name="cbox_envmap", src="pa4/cbox/cbox_path_mis.xml", res=(200, 150), rays=0, seq=1000, ref_spp=[4, 64], envmap=True),
    # BASELINE config 3: Disney + microfacet BSDFs, envmap emitter, thin-lens camera, path_mis
    dict(name="c3_project", src="pa4/cbox/cbox_path_mis.xml", res=(200, 150), rays=4000, seq=1000, ref_spp=[4, 64], envmap=True,
         camera=("thinlens", '<float name="lensRadius" value="0.03"/><float name="focalDist" value="4.9"/>'),
         swap=[('<bsdf type="mirror"/>', '<bsdf type="disney"><color name="baseColor" value="0.9,0.6,0.2"/><float name="metallic" value="0.6"/>'
                '<float name="specular" value="0.8"/><float name="specularTint" value="0.2"/><float name="roughness" value="0.3"/>'
                '<float name="sheen" value="0.3"/><float name="sheenTint" value="0.5"/></bsdf>'),
               ('<bsdf type="dielectric"/>', '<bsdf type="microfacet"><float name="alpha" value="0.15"/><color name="kd" value="0.2,0.25,0.6"/></bsdf>')]),
    # BASELINE config 5: homogeneous medium, volumetric integrator, spot light + envmap
    dict(name="c5_volumetric", src="project/volumetric/volumetric.xml", res=(200, 150), rays=0, seq=1000, ref_spp=[4, 512], envmap=True,
         extra='<emitter type="spotlight"><point name="position" value="0,1.5,0.5"/><color name="color" value="30,20,10"/>'
               '<vector name="direction" value="0.2,-1,-0.1"/><float name="falloffStart" value="15"/><float name="totalWidth" value="35"/></emitter>'),
]
# image textures, normal maps, advancedCamera (what scenes/project/final.xml uses)
FLOOR = '<bsdf type="diffuse">\n\t\t\t<color name="albedo" value=".5,.5,.5"/>\n\t\t</bsdf>'        # the table scene's floor quad
FIXTURES += [
    dict(name="table_textured", src="pa4/table/table_path_mis.xml", res=(200, 150), rays=4000, seq=1000, ref_spp=[4, 64], images=True,
         swap=[(FLOOR, '<texture type="NormalMap" name="normal"><string name="fileName" value="synth_normal.bmp"/><string name="wrap" value="clamp"/></texture>'
                '<bsdf type="diffuse"><texture type="ImageTexture" name="albedo"><string name="fileName" value="synth_albedo.bmp"/>'
                '<string name="wrap" value="repeat"/></texture></bsdf>')]),
    dict(name="cbox_advcam", src="pa4/cbox/cbox_path_mis.xml", res=(200, 150), rays=0, seq=1000, ref_spp=[4, 64],
         camera=("advancedCamera", '<float name="lensRadius" value="0.04"/><float name="focalDist" value="4.6"/>'
                 '<vector name="chromaticAberation" value="4, 2, 3.3"/><vector name="distortion" value="3, 3"/>')),
    dict(name="cbox_perlin", src="pa4/cbox/cbox_path_mis.xml", res=(200, 150), rays=6000, seq=1000, ref_spp=[4, 64],
         extra='<mesh type="perlinsphere"><point name="center" value="0.0 0.9 0.1"/><float name="radius" value="0.22"/>'
               '<float name="height" value="0.07"/><float name="scale" value="0.09"/>'
               '<bsdf type="diffuse"><color name="albedo" value="0.3 0.6 0.8"/></bsdf></mesh>'
               '<mesh type="perlinsphere"><point name="center" value="0.55 1.25 -0.3"/><float name="radius" value="0.06"/>'
               '<float name="height" value="0.03"/><float name="scale" value="0.03"/>'
               '<emitter type="area"><color name="radiance" value="25 20 12"/></emitter></mesh>'),
    dict(name="cbox_advcam_distortion", src="pa4/cbox/cbox_path_mis.xml", res=(200, 150), rays=0, seq=1000, ref_spp=[4, 64],
         camera=("advancedCamera", '<vector name="distortion" value="1.7, 1.7"/>')),
]
TTESTS = [
    dict(name="ttest_pa4_direct", src="pa4/tests/test-direct.xml"),
    dict(name="ttest_pa4_furnace", src="pa4/tests/test-furnace.xml"),
    dict(name="ttest_pa3_mesh", src="pa3/tests/test-mesh.xml"),
    dict(name="ttest_pa3_mesh_furnace", src="pa3/tests/test-mesh-furnace.xml"),
    dict(name="ttest_pa1_direct", src="pa1/test-direct.xml"),
]
# scene sources kept verbatim for the reference arm of bench.py (data files, not source code)
SCENE_COPIES = [("pa4/cbox", ["cbox_path_mis.xml", "meshes/walls.obj", "meshes/leftwall.obj", "meshes/rightwall.obj", "meshes/light.obj"])]


def run(cmd, cwd=None, timeout=3600):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed:\n{r.stdout[-2000:]}")
    return r.stdout


def synthetic_envmap(path, rows=64, cols=128, seed=5):
    """Seeded sky gradient + one bright disc (the reference's textures/*.exr are missing large blobs)."""
    rng = np.random.RandomState(seed)
    v, u = np.meshgrid(np.linspace(0, 1, cols), np.linspace(0, 1, rows))
    img = np.stack([0.3 + 0.4 * u, 0.4 + 0.4 * u, 0.9 - 0.3 * u], -1).astype(np.float32)
    img += 0.05 * rng.rand(rows, cols, 3).astype(np.float32)
    disc = ((u - 0.3) ** 2 + (v - 0.6) ** 2) < 0.004
    img[disc] = (40.0, 36.0, 30.0)
    cv2.imwrite(path, img[..., ::-1], [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_FLOAT,
                                        cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_NO])



def synthetic_images(work, w=64, h=48, seed=9):
    """Small seeded 8-bit images as 24-bit BMP (the reference's stb_image 1.4x reads PNG/JPEG/BMP/TGA): a colourful albedo and a normal map
    perturbed around +z.  Deliberately non-square so that a width/height swap cannot go unnoticed."""
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w]
    alb = np.stack([128 + 100 * np.sin(x * 0.4), 128 + 100 * np.cos(y * 0.5), 40 + 3 * ((x // 8 + y // 8) % 2) * 60], -1)
    alb = np.clip(alb + rng.randint(-10, 10, alb.shape), 0, 255).astype(np.uint8)
    n = np.stack([0.35 * np.sin(x * 0.7), 0.35 * np.cos(y * 0.9), np.ones_like(x, dtype=np.float64)], -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    nrm = np.clip((n * 0.5 + 0.5) * 255 + rng.randint(-3, 3, n.shape), 0, 255).astype(np.uint8)
    for name, img in (("synth_albedo.bmp", alb), ("synth_normal.bmp", nrm)):
        cv2.imwrite(os.path.join(work, name), np.ascontiguousarray(img[..., ::-1]))


def mirror_scene_dir(rel_dir, tmp_root):
    """Writable mirror of scenes/<rel_dir> inside tmp_root/scenes with every sibling reachable through
    symlinks (scene files use relative paths such as ../../pa1/plane.obj); returns the mirrored dir."""
    src = os.path.join(REF, "scenes")
    dst = os.path.join(tmp_root, "scenes")
    os.makedirs(dst)
    for part in rel_dir.split("/"):
        for e in os.listdir(src):
            if e != part:
                os.symlink(os.path.join(src, e), os.path.join(dst, e))
        src, dst = os.path.join(src, part), os.path.join(dst, part)
        os.makedirs(dst)
    for e in os.listdir(src):
        if e in ("ref", "images") or e.endswith((".png", ".exr")):
            continue
        os.symlink(os.path.join(src, e), os.path.join(dst, e))
    return dst

def rewrite(xml, fx, spp):
    if "res" in fx:
        xml = re.sub(r'(name="width"\s+value=")\d+', rf"\g<1>{fx['res'][0]}", xml)
        xml = re.sub(r'(name="height"\s+value=")\d+', rf"\g<1>{fx['res'][1]}", xml)
    if re.search(r'name="sampleCount"', xml):
        xml = re.sub(r'(name="sampleCount"\s+value=")\d+', rf"\g<1>{spp}", xml)
    else:
        xml = xml.replace("</scene>", f'<sampler type="independent"><integer name="sampleCount" value="{spp}"/></sampler></scene>')
    if "integrator" in fx:
        typ, body = fx["integrator"]
        xml = re.sub(r'<integrator type="[^"]*"\s*(/>|>.*?</integrator>)', f'<integrator type="{typ}">{body}</integrator>', xml, flags=re.S)
    if "camera" in fx:
        typ, body = fx["camera"]
        xml = re.sub(r'<camera type="[^"]*">', f'<camera type="{typ}">{body}', xml)
    for a, b in fx.get("swap", []):
        assert a in xml, a
        xml = xml.replace(a, b)
    if "extra" in fx:
        xml = xml.replace("</scene>", fx["extra"] + "</scene>")
    if fx.get("envmap"):
        xml = xml.replace("</scene>", '<mesh type="sphere"><point name="center" value="0,1,0"/><float name="radius" value="20"/>'
                          '<emitter type="envmap"><string name="filename" value="envmap_synth.exr"/></emitter></mesh></scene>')
        # open the box: drop the back-facing light so that the environment matters
    return xml


def make_scene_fixture(fx, tmp, meta):
    name = fx["name"]
    work = mirror_scene_dir(os.path.dirname(fx["src"]), os.path.join(tmp, name))
    if fx.get("envmap"):
        synthetic_envmap(os.path.join(work, "envmap_synth.exr"))
    if fx.get("images"):
        synthetic_images(work)
    src_xml = open(os.path.join(work, os.path.basename(fx["src"]))).read()
    entry = dict(source=fx["src"], res=fx.get("res"), ref_spp=fx["ref_spp"], rays=fx["rays"], seq=fx["seq"])
    first = True
    for spp in fx["ref_spp"]:
        xml_path = os.path.join(work, f"{name}.xml")
        open(xml_path, "w").write(rewrite(src_xml, fx, spp))
        if first:
            out = os.path.join(HERE, f"{name}.nscene")
            run([EXPORT, xml_path, out, "--rays", str(fx["rays"]), "--seq", str(fx["seq"]), "--probe", "512"], cwd=work)
            first = False
        log = run(["timeout", "3000", NORI, xml_path], cwd=work)
        m = re.search(r"took ([0-9.]+)(ms|s|m)", log)
        entry.setdefault("ref_time", {})[str(spp)] = m.group(0) if m else None
        img = cv2.imread(os.path.join(work, f"{name}.exr"), cv2.IMREAD_UNCHANGED)[..., ::-1]
        np.save(os.path.join(HERE, f"{name}.ref{spp}.npy"), np.ascontiguousarray(img, dtype=np.float32))
        if fx.get("variance") and spp == fx["ref_spp"][0]:        # the reference's second output (render.cpp:263-278)
            var = cv2.imread(os.path.join(work, f"{name}_variance.exr"), cv2.IMREAD_UNCHANGED)[..., ::-1]
            np.save(os.path.join(HERE, f"{name}.refvar{spp}.npy"), np.ascontiguousarray(var, dtype=np.float32))
    meta["scenes"][name] = entry
    print(f"[fixtures] {name}: ok", flush=True)


def make_ttest_fixture(tt, tmp, meta):
    name = tt["name"]
    work = mirror_scene_dir(os.path.dirname(tt["src"]), os.path.join(tmp, name))
    tree = ET.parse(os.path.join(work, os.path.basename(tt["src"])))
    root = tree.getroot()
    props = {c.get("name"): c.get("value") for c in root if c.tag in ("string", "float", "integer")}
    refs = [float(v) for v in re.split(r"[,\s]+", props["references"].strip()) if v]
    scenes = [c for c in root if c.tag == "scene"]
    assert len(refs) == len(scenes), (name, len(refs), len(scenes))
    entry = dict(source=tt["src"], references=refs, significance=float(props.get("significanceLevel", 0.01)),
                 sampleCount=int(props.get("sampleCount", 100000)), scenes=[])
    for i, sc in enumerate(scenes):
        xml_path = os.path.join(work, f"{name}_{i}.xml")
        ET.ElementTree(sc).write(xml_path)
        out = os.path.join(HERE, f"{name}_{i}.nscene")
        run([EXPORT, xml_path, out], cwd=work)
        entry["scenes"].append(f"{name}_{i}.nscene")
    meta["ttests"][name] = entry
    print(f"[fixtures] {name}: {len(scenes)} scenes", flush=True)


def main():
    want = set(sys.argv[1:])
    run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    meta_path = os.path.join(HERE, "meta.json")
    meta = json.load(open(meta_path)) if os.path.exists(meta_path) else {"scenes": {}, "ttests": {}}
    with tempfile.TemporaryDirectory() as tmp:
        for fx in FIXTURES:
            if not want or fx["name"] in want:
                make_scene_fixture(fx, tmp, meta)
        for tt in TTESTS:
            if not want or tt["name"] in want:
                make_ttest_fixture(tt, tmp, meta)
    if not want or "scenes" in want:
        for d, files in SCENE_COPIES:
            for f in files:
                dst = os.path.join(HERE, "scenes", os.path.basename(d), f)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(os.path.join(REF, "scenes", d, f), dst)
                os.chmod(dst, 0o644)
    # pcg32 known answers published by the reference (ext/pcg32/pcg32-demo.out:8; seed(42, 54))
    meta["pcg32_demo"] = {"initstate": 42, "initseq": 54,
                          "uint": [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]}
    json.dump(meta, open(meta_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
