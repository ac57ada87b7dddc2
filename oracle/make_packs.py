#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY — runs the reference's own XML parser (csrt::LoadConfig,
src/parser/parser.cpp:94, compiled into oracle/_ref) over the scenes BASELINE.json names and
writes them as scene packs under scenes/, so they can travel to a machine without /root/reference.

Usage: python oracle/make_packs.py [name ...]     (needs /root/reference; run once in the build container)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from refcheck import RefLib  # noqa: E402

REF_SCENES = "/root/reference/resources/scene"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes")

# name -> (xml, width, height, spp) ; width/height/spp 0 = keep the XML's value.
SCENES = {
    "cornell-box": ("cornell-box/scene_v0.6.xml", 0, 0, 0),
    "dragon": ("dragon/scene.xml", 1024, 1024, 256),                       # C2
    "mercury": ("mercury/smooth_diffuse.xml", 256, 256, 32),               # C1
    "matpreview": ("matpreview/rough_conductor.xml", 1024, 1024, 512),     # C3
    "volumetric-caustic": ("volumetric-caustic/scene_v0.6.xml", 1024, 1024, 2048),  # C4
    # the remaining scenes of resources/scene (XML's own size and spp)
    "box": ("box/scene_v0.6.xml", 0, 0, 0),
    "classroom": ("classroom/scene_v0.6.xml", 0, 0, 0),
    "dining-room": ("dining-room/scene_v0.6.xml", 0, 0, 0),
    "lte-orb-silver": ("lte-orb/silver.xml", 0, 0, 0),
    "lte-orb-rough-glass": ("lte-orb/rough_glass.xml", 0, 0, 0),
    "material-testball": ("material-testball/scene_v0.6.xml", 0, 0, 0),
}


def main():
    names = sys.argv[1:] or list(SCENES)
    os.makedirs(OUT, exist_ok=True)
    ref = RefLib("mt")
    for name in names:
        xml, w, h, spp = SCENES[name]
        dst = os.path.join(OUT, name + ".b200scene")
        ref.pack_from_xml(os.path.join(REF_SCENES, xml), dst, w, h, spp)
        print(f"{name}: {os.path.getsize(dst) / 1e6:.2f} MB -> {dst}")


if __name__ == "__main__":
    main()
