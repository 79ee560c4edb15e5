"""Bit-packing kernels on a B200 (b200_pack_codes / b200_pack_labels through the C-ABI): the 128-bit form (cols % 4 == 0,
cols <= 128, aligned source; 8 / 16 / 32 lanes per row), the scalar form (every other width, unaligned sources,
B200_PACK_V4=0) and the world-1 exchange kernels, against numpy on the same seeded inputs — bit-exact.

Reference hand-over format: float32 +-1 codes / multi-hot labels, /root/reference/main/engine/evaluate.py:26-64."""
import ctypes

import numpy as np
import pytest
import torch

from simlib import multi_hot, pack_bits, pack_labels_np, pm1, words

pytestmark = pytest.mark.gpu


def _pack(kind, host, rows, cols, offset_floats=0):
    """C-ABI call on a device copy of `host` that starts `offset_floats` floats into its allocation."""
    from image_retrieval_wavelet_b200 import _cabi

    lib = _cabi.load()
    flat = torch.zeros(rows * cols + offset_floats + 8, dtype=torch.float32, device="cuda")
    flat[offset_floats:offset_floats + rows * cols] = torch.from_numpy(np.ascontiguousarray(host, np.float32).reshape(-1)).cuda()
    src = flat[offset_floats:]
    nw = _cabi.code_words(cols) if kind == "codes" else _cabi.label_words(cols)
    out = torch.full(((rows + 1) // 2 * 2 + 2, nw), -1, dtype=torch.int64, device="cuda")       # + a canary row pair
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    fn = lib.b200_pack_codes if kind == "codes" else lib.b200_pack_labels
    _cabi.check(fn(_cabi.ptr(src), rows, cols, _cabi.ptr(out), _cabi.ptr(bad), _cabi.stream_ptr()), "pack")
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint64)
    assert (got[(rows + 1) // 2 * 2:] == np.uint64(0xFFFFFFFFFFFFFFFF)).all(), "the kernel wrote past the padded rows"
    return got[:(rows + 1) // 2 * 2], int(bad.item())


@pytest.mark.parametrize("cols", [4, 20, 24, 32, 36, 48, 64, 68, 80, 100, 128, 132, 200, 256, 1, 21, 63, 97])
@pytest.mark.parametrize("v4", ["1", "0"])
def test_pack_every_width_and_row_count(cols, v4, monkeypatch):
    monkeypatch.setenv("B200_PACK_V4", v4)
    rng = np.random.default_rng(cols)
    for rows in (1, 2, 3, 15, 16, 17, 31, 32, 33, 1000, 4097):
        c = pm1(rng, rows, cols)
        got, bad = _pack("codes", c, rows, cols)
        assert bad == 0 and np.array_equal(got, pack_bits(c, words(cols))), (rows, cols)
        lab = multi_hot(rng, rows, cols, 0.3) if cols > 1 else None
        if lab is not None:
            got, bad = _pack("labels", lab, rows, cols)
            want = pack_labels_np(lab)[0]
            assert bad == 0 and np.array_equal(got[:, :want.shape[1]], want) and not got[:, want.shape[1]:].any(), (rows, cols)


@pytest.mark.parametrize("cols", [24, 64, 80, 128])
def test_pack_from_an_unaligned_source_takes_the_scalar_kernel(cols):
    rng = np.random.default_rng(7 + cols)
    c = pm1(rng, 777, cols)
    for off in (1, 2, 3):                                        # 4, 8, 12 bytes past a 16-byte boundary
        got, bad = _pack("codes", c, 777, cols, offset_floats=off)
        assert bad == 0 and np.array_equal(got, pack_bits(c, words(cols)))


@pytest.mark.parametrize("v4", ["1", "0"])
def test_invalid_entries_are_counted_exactly(v4, monkeypatch):
    monkeypatch.setenv("B200_PACK_V4", v4)
    rng = np.random.default_rng(3)
    for cols in (24, 64, 80, 128):
        c = pm1(rng, 501, cols)
        where = rng.choice(c.size, 37, replace=False)
        c.reshape(-1)[where] = rng.choice(np.array([0.0, 0.5, -2.0, np.nan, np.inf, -np.inf, 1e-40], np.float32), 37)
        got, bad = _pack("codes", c, 501, cols)
        assert bad == 37
        assert np.array_equal(got, pack_bits(np.where(c > 0, 1.0, -1.0).astype(np.float32), words(cols)))      # bit = (x > 0), whatever x is
        lab = multi_hot(rng, 501, cols, 0.3)
        lab.reshape(-1)[where] = rng.choice(np.array([0.5, -1.0, 2.0, np.nan, np.inf, 1e-40], np.float32), 37)
        _, bad = _pack("labels", lab, 501, cols)
        assert bad == 37


def test_world_one_exchange_kernels_match_the_local_ones():
    """b200_pack_to_ranks and b200_comm_put_barrier_final on a one-rank region (no peer needed): the packed shard lands at
    its offset, and put + barrier + mean gives the very bits of b200_map_final."""
    from image_retrieval_wavelet_b200 import _cabi

    lib = _cabi.load()
    rng = np.random.default_rng(11)
    rows, bits, nq = 1501, 128, 613
    c = pm1(rng, rows, bits)
    ap = rng.random(nq)
    handle = ctypes.c_void_p()
    nbytes = 1 << 20
    _cabi.check(lib.b200_comm_create(0, 1, nbytes, ctypes.byref(handle)), "b200_comm_create")
    try:
        lib.b200_comm_buffer.restype = ctypes.c_void_p
        base = lib.b200_comm_buffer(handle, 0)
        src = torch.from_numpy(c).cuda()
        bad = torch.zeros(1, dtype=torch.int32, device="cuda")
        off_codes, off_ap, off_status = 4096, 512 * 1024, 768 * 1024
        _cabi.check(lib.b200_pack_to_ranks(_cabi.ptr(src), 1, rows, bits, handle, off_codes, _cabi.ptr(bad), _cabi.stream_ptr()),
                    "b200_pack_to_ranks")
        d_ap = torch.from_numpy(np.concatenate([ap, np.zeros(3)])).cuda()                  # 16-byte multiple
        d_status = torch.zeros(4, dtype=torch.int32, device="cuda")
        out_fused = torch.zeros(2, dtype=torch.float64, device="cuda")
        out_plain = torch.zeros(2, dtype=torch.float64, device="cuda")
        for flag in (0, 1):
            d_status[0] = flag
            srcs = (ctypes.c_void_p * 2)(d_ap.data_ptr(), d_status.data_ptr())
            offs = (ctypes.c_size_t * 2)(off_ap, off_status)
            sizes = (ctypes.c_size_t * 2)(d_ap.numel() * 8, 16)
            _cabi.check(lib.b200_comm_put_barrier_final(handle, 2, srcs, offs, sizes, off_ap, nq, off_status, _cabi.ptr(out_fused),
                                                        _cabi.stream_ptr()), "b200_comm_put_barrier_final")
            _cabi.check(lib.b200_map_final(d_ap.data_ptr(), nq, d_status.data_ptr(), 1, 16, _cabi.ptr(out_plain), _cabi.stream_ptr()),
                        "b200_map_final")
            torch.cuda.synchronize()
            assert out_fused.cpu().numpy().tobytes() == out_plain.cpu().numpy().tobytes()
            assert out_fused[1].item() == float(flag) and abs(out_fused[0].item() - ap.mean()) < 1e-12
        timed_out = ctypes.c_int()
        _cabi.check(lib.b200_comm_status(handle, ctypes.byref(timed_out)), "b200_comm_status")
        assert timed_out.value == 0 and int(bad.item()) == 0
        from image_retrieval_wavelet_b200.engine.map_engine import _RawDeviceBytes

        raw = torch.as_tensor(_RawDeviceBytes(base, nbytes), device="cuda").clone().cpu().numpy()
        got = raw[off_codes:off_codes + (rows + 1) // 2 * 2 * 16].view(np.uint64).reshape(-1, 2)
        assert np.array_equal(got, pack_bits(c, 2))
    finally:
        lib.b200_comm_destroy(handle)
