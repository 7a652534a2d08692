"""GPU tests of the library's own radix sort (csrc/sort.cu: onesweep passes with decoupled
look-back, no CUB), through the C ABI: the generic 64-bit pair sort against a stable
torch.sort, and the depth-first ordering against the generic path / the numpy oracle at sizes
that cross every tile-size and pass-count boundary of the kernels.  Bit-exact everywhere
(reference: the cub::DeviceRadixSort::SortPairs call at CS/isect_tiles.cu:252-300)."""
import ctypes
import math

import pytest
import torch

import splat_one_b200 as S
from oracle import torch_ref as O
from splat_one_b200 import wrapper
from splat_one_b200._lib import check, get_lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _native_sort(keys, vals, end_bit):
    lib = get_lib()
    n = keys.numel()
    ka, va = keys.clone(), vals.clone()
    kb, vb = torch.empty_like(ka), torch.empty_like(va)
    ws_bytes = lib.b200splat_sort_workspace_bytes(n)
    ws = torch.empty((max(ws_bytes, 1),), device=DEV, dtype=torch.uint8)
    sel = ctypes.c_int(0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    check(lib.b200splat_isect_sort(n, end_bit, p(ka), p(va), p(kb), p(vb), p(ws), ws_bytes, ctypes.byref(sel), st), lib)
    torch.cuda.synchronize()
    return (kb, vb) if sel.value else (ka, va)


@pytest.mark.parametrize("n", [1, 31, 33, 2047, 2048, 2049, 4097, 40961, (1 << 21) - 1, (1 << 21) + 5, 5_000_003])
@pytest.mark.parametrize("end_bit", [1, 8, 9, 17, 46, 64])
def test_generic_pair_sort_is_a_stable_sort_on_the_low_bits(n, end_bit):
    if n > 100_000 and end_bit not in (17, 46):
        pytest.skip("large sizes: two bit widths are enough")
    g = torch.Generator(device=DEV).manual_seed(n * 67 + end_bit)
    # sorted bits drawn from a small pool (long runs, many ties) + noise in the ignored high bits
    width = min(end_bit, 62)
    pool = torch.randint(0, 1 << width, (max(2, n // 4),), device=DEV, generator=g, dtype=torch.int64)
    keys = pool[torch.randint(0, pool.numel(), (n,), device=DEV, generator=g)]
    if end_bit < 62:
        keys = keys | (torch.randint(0, 1 << (62 - end_bit), (n,), device=DEV, generator=g, dtype=torch.int64)
                       << end_bit)
    vals = torch.arange(n, device=DEV, dtype=torch.int32)
    ks, vs = _native_sort(keys, vals, end_bit)
    # keys are non-negative, so the radix (unsigned) order of the low bits is their integer order
    sort_bits = keys if end_bit >= 62 else keys & ((1 << end_bit) - 1)
    order = torch.sort(sort_bits, stable=True)[1]
    assert torch.equal(vs.long(), order)
    assert torch.equal(ks, keys[order])


@pytest.mark.parametrize("C,N,W,H,ts", [
    (1, 300_000, 1920, 1080, 16),   # n_isects > 2^21: 16 items per thread, two 16-bit-key passes
    (1, 3000, 100, 60, 16),         # cam|tile in 7 bits: ONE pass, which is also the final one
    (40, 3000, 1024, 1024, 16),     # 13 + 6 = 19 key bits: 32-bit keys, three passes
    (3, 70_000, 640, 360, 8),       # n_elems between the 8- and 16-item variants
])
def test_depth_first_order_equals_generic_sort_and_oracle(C, N, W, H, ts, monkeypatch):
    g = torch.Generator().manual_seed(C * 31 + N)
    m2 = torch.rand(C, N, 2, generator=g) * torch.tensor([W * 1.1, H * 1.1]) - torch.tensor([W * 0.05, H * 0.05])
    radii = torch.randint(0, 48, (C, N), generator=g, dtype=torch.int32)
    depths = torch.rand(C, N, generator=g) * 20 + 0.01
    depths[:, : N // 8] = 2.25  # ties: stability
    tw, th = math.ceil(W / ts), math.ceil(H / ts)
    args = (m2.to(DEV), radii.to(DEV), depths.to(DEV), ts, tw, th)
    fast = wrapper.isect_tiles_and_offsets(*args)
    monkeypatch.setattr(wrapper, "_FORCE_GENERIC_SORT", True)
    slow = wrapper.isect_tiles_and_offsets(*args)
    monkeypatch.setattr(wrapper, "_FORCE_GENERIC_SORT", False)
    assert fast[1].numel() == slow[1].numel() and fast[1].numel() > 0
    for a, b in zip(fast, slow):
        assert torch.equal(a, b)
    # sortedness of the 64-bit ids on the sorted bits, and offsets consistent with them
    ids = fast[1]
    assert (ids[1:] >= ids[:-1]).all()
    hi = ids >> 32
    tile_n_bits = int(tw * th).bit_length()
    lin = (hi >> tile_n_bits) * (tw * th) + (hi & ((1 << tile_n_bits) - 1))
    ref_offs = torch.searchsorted(lin, torch.arange(C * tw * th, device=DEV)).int().view(C, th, tw)
    assert torch.equal(fast[3], ref_offs)
    if C * N <= 250_000:  # the numpy oracle enumerates pairs in Python-free numpy; keep it in seconds
        ref = O.isect_tiles(m2, radii, depths, ts, tw, th)
        for a, b in zip(fast[:3], ref):
            assert torch.equal(a.cpu(), b)


def test_depth_first_order_is_deterministic_under_stream_concurrency():
    """Two streams sort different inputs at once (tickets / look-back words are per call)."""
    outs = []
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    ins = []
    for s in range(2):
        g = torch.Generator().manual_seed(900 + s)
        N = 150_000
        ins.append((torch.rand(1, N, 2, generator=g).mul(torch.tensor([640.0, 360.0])).to(DEV),
                    torch.randint(0, 30, (1, N), generator=g, dtype=torch.int32).to(DEV),
                    (torch.rand(1, N, generator=g) * 9 + 0.1).to(DEV)))
    torch.cuda.synchronize()
    for s in range(2):
        with torch.cuda.stream(streams[s]):
            outs.append(wrapper.isect_tiles_and_offsets(*ins[s], 16, 40, 23))
    torch.cuda.synchronize()
    for s in range(2):
        again = wrapper.isect_tiles_and_offsets(*ins[s], 16, 40, 23)
        for a, b in zip(outs[s], again):
            assert torch.equal(a, b)
