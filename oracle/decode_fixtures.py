"""TEST INFRASTRUCTURE: decode the reference's JPEG fixtures with the *bundled* OpenCV 2.4.13 highgui
(other JPEG decoders differ by a few LSBs) into raw files under oracle/_ref/fixtures/
named NAME_WxHxC.bin.  Run by `make -C oracle ref`; needs /root/reference."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from oracle_lib import Ref  # noqa: E402

FIXTURES = {  # name -> (path under the reference tree, colour flag)
    "universe1920x1080": ("res/universe19201080.jpeg", 1),     # config 1
    "lakers2560x1440": ("res/lakers25601440.jpeg", 1),         # config 4 shape; test_resize/test_crop fixture
    "face1280x720": ("res/face1280720.jpg", 1),                # config 3 shape
    "t1280x720": ("src/test/res/1280x720.jpg", 1),             # test_warp_affine.cpp
    "t1280x720_grey": ("src/test/res/1280x720_grey.jpg", 0),   # test_warp_affine.cpp (rotation variant)
    "t640x360": ("src/test/res/640x360.jpg", 1),
    "t284x214": ("src/test/res/284x214.jpg", 1),               # test_normalize.cpp
    "t176x144": ("src/test/res/176x144.jpg", 1),               # test_change_dtype/layout/normalize
}


def main(ref_root, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ref = Ref()
    for name, (rel, color) in FIXTURES.items():
        img = ref.imread(os.path.join(ref_root, rel), color)
        h, w, c = img.shape
        img.tofile(os.path.join(out_dir, f"{name}_{w}x{h}x{c}.bin"))
        print(f"{name}: {w}x{h}x{c}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
