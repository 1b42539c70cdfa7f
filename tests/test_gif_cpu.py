"""GIF sweep, CPU part: the library's host coder / container writer against the oracle's restatement of
GifGenerator.cs, and BOTH against an independent GIF decoder (PIL) — the only externally pinned check this repo can make,
because the reference ships no GIF fixtures.  No CUDA device is needed here: the palette mapping (a device kernel, no host
fallback) is covered by tests/test_gpu_gif.py; these tests feed the library palette indices.
"""
import importlib
import io

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

gif = importlib.import_module("cosig-raytracing_b200.gif_generator")


def _frames(n, h, w, seed=0, smooth=False):
    rng = np.random.RandomState(seed)
    if smooth:  # gradients: long LZW matches, like rendered frames
        y, x = np.mgrid[0:h, 0:w]
        out = np.stack([np.stack([(x * 255 // max(1, w - 1) + 7 * k) % 256, (y * 255 // max(1, h - 1)) % 256, ((x + y + 13 * k) // 3) % 256,
                                  np.full_like(x, 255)], axis=-1) for k in range(n)]).astype(np.uint8)
        return out
    f = rng.randint(0, 256, size=(n, h, w, 4)).astype(np.uint8)
    f[..., 3] = 255
    return f


def test_color_table(pkg, oracle):
    t = gif.color_table()
    assert (t.reshape(-1) == oracle.gif_color_table()).all()
    assert tuple(t[0]) == (0, 0, 0) and tuple(t[215]) == (255, 255, 255)          # 6x6x6 cube, GifGenerator.cs:224-235
    assert tuple(t[1]) == (0, 0, 51) and tuple(t[36]) == (51, 0, 0)
    assert tuple(t[216]) == (0, 0, 0) and tuple(t[217]) == (6, 6, 6) and tuple(t[255]) == (253, 253, 253)  # (byte)((i-216)*6.5f)


def test_convert_to_indexed_oracle_known_values(oracle):
    # one row per case; row 0 = bottom must come out last (vertical flip, GifGenerator.cs:360-366)
    px = np.zeros((2, 3, 4), np.uint8)
    px[0, 0] = (255, 255, 255, 255)   # 5,5,5 -> 215
    px[0, 1] = (42, 43, 0, 255)       # 42/255*5.99 = 0.9866 -> 0 ; 43/255*5.99 = 1.0101 -> 1
    px[0, 2] = (0, 0, 213, 255)       # 213/255*5.99 = 5.0034 -> 5
    px[1, 0] = (128, 0, 0, 255)       # 3.0067 -> 3 -> 108
    idx = oracle.gif_convert_to_indexed(px)
    assert idx.tolist() == [[108, 0, 0], [215, 6, 5]]


@pytest.mark.parametrize("data", [b"", b"\x00", b"\x07" * 1000, bytes(range(256)) * 3, bytes([1, 2] * 500)])
def test_lzw_small_cases(pkg, oracle, data):
    arr = np.frombuffer(data, np.uint8)
    assert gif.lzw_compress(arr) == oracle.gif_lzw(arr)


def test_lzw_dictionary_freezes_at_4096(pkg, oracle):
    # random bytes create a new code per ~2 input bytes: the 4096-code limit is reached and, as in the reference (:471),
    # the dictionary is simply frozen — no clear code is ever emitted again
    rng = np.random.RandomState(1)
    arr = rng.randint(0, 216, size=60000).astype(np.uint8)
    a, b = gif.lzw_compress(arr), oracle.gif_lzw(arr)
    assert a == b
    assert len(a) > 60000  # 12-bit codes for mostly single bytes: "compression" expands, bound must hold
    lib = importlib.import_module("cosig-raytracing_b200.abi").load()
    assert len(a) <= lib.rtb_gif_lzw_bound(arr.size)


@settings(max_examples=60, deadline=None)
@given(st.binary(min_size=0, max_size=3000), st.integers(1, 4))
def test_lzw_property(pkg, oracle, blob, alphabet_bits):
    arr = np.frombuffer(blob, np.uint8) & ((1 << (2 * alphabet_bits)) - 1)  # small alphabets give long matches
    assert gif.lzw_compress(arr) == oracle.gif_lzw(arr)


@pytest.mark.parametrize("shape,smooth", [((3, 23, 37), False), ((2, 48, 64), True), ((1, 1, 1), False), ((4, 30, 52), True)])
def test_file_matches_oracle_and_decodes(pkg, oracle, tmp_path, shape, smooth):
    from PIL import Image
    n, h, w = shape
    frames = _frames(n, h, w, seed=n * 100 + w, smooth=smooth)
    ours, theirs = str(tmp_path / "ours.gif"), str(tmp_path / "oracle.gif")
    gif.save_indexed(ours, [oracle.gif_convert_to_indexed(f) for f in frames], frameDelay=10, threads=2)
    oracle.gif_save(theirs, frames, 10)
    a, b = open(ours, "rb").read(), open(theirs, "rb").read()
    assert a == b, "library GIF differs from the restated GifGenerator.SaveGif output"
    # independent decoder
    im = Image.open(io.BytesIO(a))
    assert im.format == "GIF" and im.size == (w, h)
    assert getattr(im, "n_frames", 1) == n
    assert im.info.get("loop", None) == 0            # Netscape extension, :203-213
    table = oracle.gif_color_table()
    for k in range(n):
        im.seek(k)
        assert im.info.get("duration") == 100        # 10 cs
        expected = oracle.gif_convert_to_indexed(frames[k])
        decoded_rgb = np.asarray(im.convert("RGB"))
        assert (decoded_rgb == table.reshape(256, 3)[expected]).all(), f"frame {k}: PIL decodes other pixels than were encoded"


def test_save_indexed_is_independent_of_thread_count(pkg, oracle, tmp_path):
    frames = _frames(3, 20, 28, seed=5)
    indexed = [oracle.gif_convert_to_indexed(f) for f in frames]
    a, b = str(tmp_path / "a.gif"), str(tmp_path / "b.gif")
    gif.save_indexed(a, indexed, frameDelay=7, threads=1)
    gif.save_indexed(b, indexed, frameDelay=7, threads=3)
    assert open(a, "rb").read() == open(b, "rb").read()


def test_rgba_frames_need_a_context(pkg, tmp_path):
    # palette mapping is a CUDA kernel: without a context the call is refused instead of falling back to the host
    import ctypes as C
    lib = importlib.import_module("cosig-raytracing_b200.abi").load()
    frame = _frames(1, 4, 4)[0]
    ptrs = (C.c_void_p * 1)(frame.ctypes.data)
    assert lib.rtb_gif_save(None, str(tmp_path / "x.gif").encode(), 4, 4, ptrs, 1, 10, 1) == -1  # RTB_E_ARG
    assert not (tmp_path / "x.gif").exists()


def test_save_gif_empty_is_a_no_op(pkg, tmp_path):
    # GifGenerator.cs:84,162: null / empty frame lists return without touching the file system
    class Dummy:
        _ctx = None
    g = gif.GifGenerator(Dummy(), None)
    g.SaveGif([], str(tmp_path / "none.gif"))
    g.SaveGifAsync([], str(tmp_path / "none.gif"))
    assert not (tmp_path / "none.gif").exists()


def test_golden_rotation_gif(pkg, oracle, tmp_path):
    """tests/golden/rotation_test_scene_1_48x36.gif (make_golden_gif.py): the oracle still writes these bytes, the library's coder
    writes them from the oracle's indices, and PIL reads 4 frames of 48x36 out of them."""
    import os
    import sys
    from PIL import Image
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_golden_gif as G
    want = open(G.PATH, "rb").read()
    frames = G.oracle_frames()
    again = str(tmp_path / "oracle.gif")
    oracle.gif_save(again, frames, 10)
    assert open(again, "rb").read() == want, "the oracle no longer reproduces its golden GIF"
    ours = str(tmp_path / "ours.gif")
    gif.save_indexed(ours, [oracle.gif_convert_to_indexed(f) for f in frames], frameDelay=10)
    assert open(ours, "rb").read() == want
    im = Image.open(G.PATH)
    assert im.n_frames == 4 and im.size == (48, 36)
    assert len({np.asarray(im.convert("RGB")).tobytes() for _ in [im.seek(k) for k in range(4)]}) >= 1
