#!/usr/bin/env python
"""Re-pack the reference's bundled scene ASSETS (data, not source) as zip archives
under scenes/, so the benchmark inputs BASELINE.json names travel to the GPU box,
where /root/reference does not exist.

Run in the build container:  python tools/pack_scenes.py [/root/reference/resources]

Contents per archive: the .obj, its .mtl and the texture files the .mtl names,
byte-for-byte.  chinese-building is skipped: its .obj is a stripped large blob
(.MISSING_LARGE_BLOBS) so it cannot be rendered.  test_room.blend (Blender
source, unused by the renderer) is not packed.
"""
import os
import sys
import zipfile

SCENES = {
    "cube": ["cube/cube.obj", "cube/cube.mtl", "cube/cube-diffuse.jpg", "cube/cube-normal.png"],
    "teapot": ["teapot/teapot.obj", "teapot/default.mtl", "teapot/default.png"],
    "test_room": ["test_room/test_room.obj", "test_room/test_room.mtl"],
    "sonic": ["sonic.obj"],
    "living_room": None,  # whole directory
}


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/resources"
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes")
    os.makedirs(out_dir, exist_ok=True)
    for name, files in SCENES.items():
        if files is None:
            files = []
            for d, _, fs in os.walk(os.path.join(root, name)):
                for f in sorted(fs):
                    files.append(os.path.relpath(os.path.join(d, f), root))
        path = os.path.join(out_dir, name + ".zip")
        with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED, compresslevel=9) as z:
            for f in sorted(files):
                zi = zipfile.ZipInfo(f, date_time=(2024, 1, 1, 0, 0, 0))  # deterministic archives
                zi.compress_type = zipfile.ZIP_DEFLATED
                with open(os.path.join(root, f), "rb") as fh:
                    z.writestr(zi, fh.read())
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
