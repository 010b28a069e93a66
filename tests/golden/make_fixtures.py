"""Generates the golden fixtures under tests/golden/ from the reference's bundled files.

Run in the build container (needs /root/reference, which does NOT exist on the GPU box):

    python tests/golden/make_fixtures.py

What it writes
  lena_grey_256.u8      LenaGrey.png decoded (palette PNG, r=g=b) -> 256x256 bytes
  lena64.u8             Lena64.png decoded (RGBA, r=g=b)          -> 64x64 bytes
  lena_colored_256.rgb  LenaColored.jpg decoded (4:4:4 JPEG)      -> 256x256x3 bytes, RGB interleaved
  unknown_run.bin       the reference's own encoded stream of LenaColored.jpg (B=8, wk=2): the
                        byte-exact known answer of the RGB encode path
  golden.json           the five avgError labels visible in the reference's Animation.gif
                        (grey encode -> quantise -> decode of LenaGrey.png), plus sha256 digests
                        of oracle-produced streams for the parity configs, so that a later
                        change of the oracle cannot go unnoticed
  lena_grey_b8_wk2.run, lena64_b8_full.run, lena64_b4_full.run
                        oracle-produced streams (small) used as GPU-side known answers
The pixel arrays are data decoded from the reference's image files, not source code.
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def argb(rgb):
    a = rgb.astype(np.uint32)
    return (0xFF000000 | (a[..., 0] << 16) | (a[..., 1] << 8) | a[..., 2]).astype(np.uint32).view(np.int32)


def main():
    grey = np.asarray(Image.open(f"{REF}/LenaGrey.png").convert("RGB"))
    l64 = np.asarray(Image.open(f"{REF}/Lena64.png").convert("RGB"))
    col = np.asarray(Image.open(f"{REF}/LenaColored.jpg").convert("RGB"))
    assert (grey[..., 0] == grey[..., 1]).all() and (grey[..., 1] == grey[..., 2]).all()
    assert (l64[..., 0] == l64[..., 1]).all() and (l64[..., 1] == l64[..., 2]).all()
    grey[..., 0].astype(np.uint8).tofile(f"{HERE}/lena_grey_256.u8")
    l64[..., 0].astype(np.uint8).tofile(f"{HERE}/lena64.u8")
    col.astype(np.uint8).tofile(f"{HERE}/lena_colored_256.rgb")
    with open(f"{REF}/unknown.run", "rb") as f:
        ref_stream = f.read()
    with open(f"{HERE}/unknown_run.bin", "wb") as f:
        f.write(ref_stream)

    gold = {
        "source": "LariWa/Fractal-Image-Compression: Animation.gif labels (RLEAppController.java:180), unknown.run",
        "gif_avg_error": [  # (B, wk, label) -- frames listed in SURVEY.md section 4
            [16, 16, "0.3744049"], [8, 16, "0.3647766"], [4, 16, "0.73760986"],
            [8, 8, "0.52404785"], [8, 4, "0.36376953"],
        ],
        "unknown_run_sha256": hashlib.sha256(ref_stream).hexdigest(),
        "oracle_streams": {},
    }
    G, L, Cc = argb(grey), argb(l64), argb(col)
    cases = {
        "lena_grey_b8_wk2": (G, 8, 2, False), "lena_grey_b8_wk16": (G, 8, 16, False),
        "lena_grey_b16_wk16": (G, 16, 16, False), "lena_grey_b4_wk16": (G, 4, 16, False),
        "lena_grey_b8_full": (G, 8, 61, False), "lena_grey_b16_full": (G, 16, 29, False),
        "lena_grey_b4_wk8": (G, 4, 8, False),
        "lena64_b8_wk2": (L, 8, 2, False), "lena64_b8_full": (L, 8, 13, False), "lena64_b4_full": (L, 4, 29, False),
        "lena_colored_b8_wk2": (Cc, 8, 2, True), "lena_colored_b4_wk4": (Cc, 4, 4, True),
        "lena_colored_b16_wk2": (Cc, 16, 2, True),
    }
    for name, (img, B, wk, rgb) in cases.items():
        H, W = img.shape
        info = O.encode(img, B, wk, rgb=rgb, nthreads=8)
        s = O.write_data(info, W, H, B, wk, rgb=rgb)
        dec, avg, it = O.decode(s)
        gold["oracle_streams"][name] = {
            "B": B, "wk": wk, "rgb": rgb, "W": W, "H": H,
            "stream_sha256": hashlib.sha256(s).hexdigest(),
            "decoded_sha256": hashlib.sha256(dec.tobytes()).hexdigest(),
            "avg_error": repr(float(avg)), "avg_error_hex": float(avg).hex(), "iterations": it,
        }
        if name in ("lena_grey_b8_wk2", "lena64_b8_full", "lena64_b4_full"):
            with open(f"{HERE}/{name}.run", "wb") as f:
                f.write(s)
    assert O.write_data(O.encode(Cc, 8, 2, rgb=True), 256, 256, 8, 2, rgb=True) == ref_stream
    with open(f"{HERE}/golden.json", "w") as f:
        json.dump(gold, f, indent=1)
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
