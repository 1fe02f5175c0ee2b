"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute calls without a GPU), ops refuse CPU tensors, and the host-side mirrors of
the reference classes behave like the reference (known-answer vectors re-stated from
tests/test_masks.py of the reference)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import dynamask_b200 as dm
from dynamask_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'dynamask_sm100.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dm_[a-z_0-9]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert {'dm_assign', 'dm_roi_align_fwd', 'dm_roi_align_bwd', 'dm_paste_masks', 'dm_mask_target'} <= set(syms)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), s
    assert set(_lib.SIGNATURES) == set(syms)


def test_library_version_and_error_strings():
    lib = _lib.load()
    assert lib.dm_version() >= 100
    assert lib.dm_error_string(0) == b'ok'
    assert b'invalid' in lib.dm_error_string(-1)
    assert lib.dm_launch_count() >= 0


def test_ops_refuse_cpu_tensors():
    f = torch.randn(1, 2, 8, 8)
    r = torch.tensor([[0, 1.0, 1.0, 5.0, 5.0]])
    with pytest.raises(NotImplementedError):
        dm.roi_align(f, r, 2)
    with pytest.raises(NotImplementedError):
        ops.assign(r, None, 4, 56.0, 1)
    with pytest.raises(NotImplementedError):
        ops.paste_masks(torch.rand(1, 1, 4, 4), torch.tensor([[0, 0, 4.0, 4.0]]), None, 8, 8, [0, 0, 8, 8], False, 0.5, 0)
    bm = dm.BitmapMasks(np.ones((1, 8, 8), np.uint8), 8, 8)
    with pytest.raises(NotImplementedError):
        bm.crop_and_resize(np.array([[0, 0, 4, 4]], np.float32), (4, 4), np.array([0]), device='cpu')


def test_roi_align_module_surface():
    layer = dm.RoIAlign(14, spatial_scale=0.25, sampling_ratio=0)
    assert layer.output_size == (14, 14) and layer.aligned is True and layer.use_torchvision is False
    assert layer.pool_mode == 'avg' and 'RoIAlign(output_size=(14, 14)' in repr(layer)
    with pytest.raises(NotImplementedError):
        dm.RoIAlign(7, pool_mode='max')
    ext = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=2), 256, [4, 8, 16, 32])
    assert ext.num_inputs == 4 and len(ext.roi_layers) == 4 and ext.finest_scale == 56
    assert [l.spatial_scale for l in ext.roi_layers] == [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    assert ext.roi_layers[0].sampling_ratio == 2 and list(ext.parameters()) == []
    with pytest.raises(AssertionError):
        dm.SingleRoIExtractor(dict(type='NoSuchLayer', output_size=7), 256, [4])


def test_roi_rescale_and_bbox2roi():
    ext = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=0), 8, [4])
    rois = torch.tensor([[1.0, 10.0, 20.0, 30.0, 60.0]])
    out = ext.roi_rescale(rois, 2.0)
    assert torch.allclose(out, torch.tensor([[1.0, 0.0, 0.0, 40.0, 80.0]]))
    r = dm.bbox2roi([torch.tensor([[1.0, 2.0, 3.0, 4.0, 0.9]]), torch.zeros(0, 5), torch.tensor([[5.0, 6.0, 7.0, 8.0, 0.5]])])
    assert torch.equal(r, torch.tensor([[0.0, 1, 2, 3, 4], [2.0, 5, 6, 7, 8]]))


def test_force_fp32_guard():
    from dynamask_b200.fp16_utils import force_fp32

    class M(torch.nn.Module):
        fp16_enabled = False

        @force_fp32(apply_to=('feats', ), out_fp16=True)
        def forward(self, feats, rois):
            return feats[0] * 2, feats[0].dtype, rois.dtype

    m = M()
    x = [torch.ones(2, dtype=torch.half)]
    r = torch.ones(1, dtype=torch.half)
    assert m(x, r)[1] == torch.half
    m.fp16_enabled = True
    y, seen, seen_r = m(x, r)
    assert seen == torch.float32 and seen_r == torch.half and y.dtype == torch.half


def test_gumbel_softmax_matches_oracle():
    from oracle import oracle as O
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(32, 4, generator=g, requires_grad=True)
    u = torch.rand(32, 4, generator=g)
    noise = -torch.log(-torch.log(u + 1e-20) + 1e-20)
    y = dm.get_mask_label(logits, noise=noise)
    hard, ind = O.gumbel_softmax_hard(logits.detach(), u)
    assert torch.equal(y.detach().argmax(1), ind)
    assert torch.allclose(y.detach(), hard, atol=1e-6)
    y.sum().backward()            # straight-through: gradient flows through the soft sample
    assert logits.grad is not None


# ---- BitmapMasks known-answer behaviour (vectors re-stated from the reference's tests/test_masks.py) ----
def test_bitmapmasks_container_and_transforms():
    raw = np.random.RandomState(0).randint(0, 2, (3, 28, 28), dtype=np.uint8)
    bm = dm.BitmapMasks(raw, 28, 28)
    assert len(bm) == 3 and bm.height == 28 and bm.width == 28
    assert (bm.to_ndarray() == raw).all()
    assert bm[1].masks.shape == (1, 28, 28) and bm[[0, 2]].masks.shape == (2, 28, 28)
    assert repr(bm) == 'BitmapMasks(num_masks=3, height=28, width=28)'
    # flip is an involution
    assert (bm.flip('horizontal').flip('horizontal').masks == raw).all()
    assert (bm.flip('vertical').masks == raw[:, ::-1]).all()
    with pytest.raises(AssertionError):
        bm.flip('diagonal')
    # pad adds zeros
    padded = bm.pad((56, 56))
    assert padded.masks.shape == (3, 56, 56) and (padded.masks[:, 28:, 28:] == 0).all()
    assert (padded.masks[:, :28, :28] == raw).all()
    # expand
    ex = bm.expand(40, 44, 5, 7)
    assert ex.masks.shape == (3, 40, 44) and (ex.masks[:, 5:33, 7:35] == raw).all() and ex.masks.sum() == raw.sum()
    # areas
    assert (bm.areas == raw.sum((1, 2))).all()
    # to_tensor
    t = bm.to_tensor(torch.uint8, 'cpu')
    assert t.shape == (3, 28, 28) and (t.numpy() == raw).all()
    # empty
    empty = dm.BitmapMasks(np.zeros((0, 28, 28), np.uint8), 28, 28)
    assert len(empty) == 0 and empty.pad((56, 56)).masks.shape == (0, 56, 56)
    assert empty.crop_and_resize(np.zeros((0, 4), np.float32), (14, 14), np.zeros(0, np.int64), device='cuda').masks.shape == (0, 14, 14)


def test_bitmapmasks_rescale_resize_crop_known_answers():
    # tests/test_masks.py:86-95 (rescale), :107-118 (resize), :170-188 (crop)
    raw = np.array([[[1, 0, 0, 0], [0, 1, 0, 1]]]).astype(np.uint8)
    bm = dm.BitmapMasks(raw, 2, 4)
    rs = bm.rescale((8, 8))
    assert rs.height == 4 and rs.width == 8
    truth = np.array([[[1, 1, 0, 0, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0, 0, 0],
                       [0, 0, 1, 1, 0, 0, 1, 1], [0, 0, 1, 1, 0, 0, 1, 1]]])
    assert (rs.to_ndarray() == truth).all()
    raw = np.diag(np.ones(4, dtype=np.uint8))[np.newaxis, ...]
    bm = dm.BitmapMasks(raw, 4, 4)
    rz = bm.resize((8, 8))
    truth = np.array([[[1, 1, 0, 0, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1, 0, 0, 0, 0],
                       [0, 0, 1, 1, 0, 0, 0, 0], [0, 0, 0, 0, 1, 1, 0, 0], [0, 0, 0, 0, 1, 1, 0, 0],
                       [0, 0, 0, 0, 0, 0, 1, 1], [0, 0, 0, 0, 0, 0, 1, 1]]])
    assert (rz.to_ndarray() == truth).all() and rz.height == 8
    rz2 = bm.resize((4, 8))
    truth = np.array([[[1, 1, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1, 0, 0, 0, 0], [0, 0, 0, 0, 1, 1, 0, 0],
                       [0, 0, 0, 0, 0, 0, 1, 1]]])
    assert (rz2.to_ndarray() == truth).all()
    raw = np.random.RandomState(1).randint(0, 2, (3, 28, 28), dtype=np.uint8)
    bm = dm.BitmapMasks(raw, 28, 28)
    cr = bm.crop(np.array([0, 10, 10, 27], dtype=np.int64))
    assert cr.masks.shape == (3, 17, 10) and (cr.masks == raw[:, 10:27, 0:10]).all()
    with pytest.raises(AssertionError):
        bm.crop(np.array([[0, 0, 4, 4]]))


def test_mask_target_requires_cuda_and_bitmaps():
    class Cfg:
        mask_size = 14
    with pytest.raises(NotImplementedError):
        dm.mask_target([torch.zeros(1, 4)], [torch.zeros(1, dtype=torch.long)],
                       [dm.BitmapMasks(np.ones((1, 8, 8), np.uint8), 8, 8)], Cfg)
    assert dm.mask_target([], [], [], Cfg) == []


def test_pack_bitmaps_layout():
    from dynamask_b200.mask_structures import pack_bitmaps
    a = np.arange(2 * 3 * 4, dtype=np.uint8).reshape(2, 3, 4)
    b = np.arange(1 * 2 * 5, dtype=np.uint8).reshape(1, 2, 5) + 100
    blob, offs, ghw = pack_bitmaps([a, b], 'cpu')
    assert blob.dtype == torch.uint8 and offs.tolist() == [0, 24]
    assert ghw.view(-1, 3).tolist() == [[2, 3, 4], [1, 2, 5]]
    assert blob[:24].tolist() == a.reshape(-1).tolist() and blob[24:34].tolist() == b.reshape(-1).tolist()


def test_sharding_helpers():
    assert dm.image_shard(10, 1, 4) == [1, 5, 9]
    rois = torch.tensor([[0.0, 1, 1, 2, 2], [1.0, 3, 3, 4, 4], [2.0, 5, 5, 6, 6], [3.0, 7, 7, 8, 8], [1.0, 9, 9, 10, 10]])
    local, idx = dm.shard_rois(rois, 1, 2)
    assert idx.tolist() == [1, 3, 4] and local[:, 0].tolist() == [0.0, 1.0, 0.0]
    a = torch.arange(12, dtype=torch.float32)
    assert dm.checksum64(a) == dm.checksum64(a.clone()) != dm.checksum64(a.flip(0))
    assert dm.gather_checksums([1, 2]) == [[1, 2]]


# ------------------------------------------------------------------------------------------
# COCO RLE (next row): the host half of the path runs without a GPU
# ------------------------------------------------------------------------------------------
def _transitions(mask):
    """What the kernels emit: every column on its own, from 0 and back to 0 after its last row."""
    h, w = mask.shape
    out = []
    for x in range(w):
        col = np.concatenate([[0], mask[:, x].astype(np.int64), [0]])
        out.extend((x * h + np.flatnonzero(col[1:] != col[:-1])).tolist())
    return np.asarray(out, np.int32)


def _compress(trans, total):
    lib = _lib.load()
    buf = ctypes.create_string_buffer(6 * (len(trans) + 1) + 8)
    t = np.ascontiguousarray(trans, np.int32)
    n = lib.dm_rle_compress_host(ctypes.c_void_p(t.ctypes.data), int(t.size), int(total),
                                 ctypes.cast(buf, ctypes.c_void_p), len(buf))
    assert n >= 0
    return buf.raw[:n]


def test_oracle_rle_round_trip_and_small_known_cases():
    from oracle import oracle as O
    assert O.rle_counts(np.array([[0, 1], [1, 1]])) == [1, 3]          # column-major 0,1,1,1
    assert O.rle_counts(np.ones((3, 3))) == [0, 9]                      # starts with an empty zero run
    assert O.rle_counts(np.zeros((5, 7))) == [35]
    assert O.rle_encode(np.array([[0, 1], [1, 1]]))['counts'] == b'13'
    rng = np.random.default_rng(3)
    for _ in range(100):
        h, w = int(rng.integers(1, 50)), int(rng.integers(1, 50))
        m = (rng.random((h, w)) < rng.random()).astype(np.uint8)
        r = O.rle_encode(m)
        assert r['size'] == [h, w]
        assert O.rle_from_string(r['counts']) == O.rle_counts(m)
        assert np.array_equal(O.rle_decode(r), m)


def test_rle_compress_host_matches_oracle_strings():
    from oracle import oracle as O
    rng = np.random.default_rng(4)
    cases = [np.zeros((6, 9), np.uint8), np.ones((6, 9), np.uint8)]
    full = np.zeros((8, 5), np.uint8)
    full[:, 1:4] = 1                      # runs that continue across column boundaries (pairs cancel)
    cases.append(full)
    big = np.zeros((800, 1333), np.uint8)
    big[100:300, 200:500] = 1             # counts far above 2^15: multi-character varints, negative deltas
    big[0:800, 900] = 1
    cases.append(big)
    for _ in range(60):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        cases.append((rng.random((h, w)) < rng.random()).astype(np.uint8))
    for m in cases:
        got = _compress(_transitions(m), m.size)
        assert got == O.rle_encode(m)['counts'], m.shape
    # too small a buffer is reported, not overrun
    lib = _lib.load()
    t = _transitions(big)
    buf = ctypes.create_string_buffer(4)
    assert lib.dm_rle_compress_host(ctypes.c_void_p(t.ctypes.data), int(t.size), int(big.size),
                                    ctypes.cast(buf, ctypes.c_void_p), 4) == -1


def test_rle_ops_refuse_cpu_tensors():
    with pytest.raises(NotImplementedError):
        ops.rle_from_canvas(torch.zeros(1, 4, 4, dtype=torch.bool))
    with pytest.raises(NotImplementedError):
        ops.paste_rle(torch.zeros(1, 1, 4, 4), torch.zeros(1, 4), None, 8, 8, [0, 0, 8, 8], True, 0.5)


def test_ops_call_keeps_the_dispatcher_for_cpu_tensors_and_autograd():
    """ops.call runs a custom op's Python body directly only for plain CUDA tensors with nothing to differentiate;
    CPU tensors still reach the dispatcher (and raise: there is no CPU implementation), plain functions are called."""
    assert ops.call(lambda a, b: a + b, 2, 3) == 5
    rois = torch.zeros(3, 5)
    with pytest.raises(NotImplementedError):
        ops.call(ops.assign, rois, None, 4, 56.0, 1)
    import dynamask_b200 as dm
    ext = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=0), 4, [4, 8])
    with pytest.raises(NotImplementedError):
        ext([torch.zeros(1, 4, 16, 16), torch.zeros(1, 4, 8, 8)], rois)
    with torch.no_grad(), pytest.raises(NotImplementedError):
        ext([torch.zeros(1, 4, 16, 16), torch.zeros(1, 4, 8, 8)], rois)


def test_paste_rle_strings_workspace_is_declared_and_monotone():
    lib = _lib.load()
    a = lib.dm_paste_rle_strings_workspace(100, 1333, 800, 1 << 16)
    b = lib.dm_paste_rle_strings_workspace(100, 1333, 800, 1 << 17)
    c = lib.dm_paste_rle_strings_workspace(200, 1333, 800, 1 << 16)
    assert 0 < a < b and a < c and a % 16 == 0
    assert lib.dm_paste_rle_strings_workspace(-1, 1, 1, 1) == -1


def test_polygon_masks_container_and_transforms():
    """Host side of PolygonMasks (reference tests/test_masks.py:300-330, :413-470 minus the
    rasterisation, which needs the device)."""
    import dynamask_b200 as dm
    PM = dm.PolygonMasks
    with pytest.raises(AssertionError):
        PM(np.zeros((3, 28, 28)), 28, 28)            # not a list
    raw = [[np.array([1., 1, 3, 1, 4, 3, 2, 4, 1, 3])], [np.array([0., 0, 1, 0, 1, 1]), np.array([1., 1, 2, 1, 2, 2, 1, 2])]]
    pm = PM(raw, 5, 5)
    assert len(pm) == 2 and repr(pm) == 'PolygonMasks(num_masks=2, height=5, width=5)'
    assert len(pm[0]) == 1 and len(pm[[1, 0]]) == 2 and len(pm[np.array([1])]) == 1
    rs = pm.resize((10, 20))
    assert (rs.height, rs.width) == (10, 20)
    assert np.allclose(rs.masks[0][0][0::2], raw[0][0][0::2] * 4) and np.allclose(rs.masks[0][0][1::2], raw[0][0][1::2] * 2)
    assert np.allclose(pm.flip('horizontal').masks[0][0][0::2], 5 - raw[0][0][0::2])
    assert np.allclose(pm.flip('vertical').flip('vertical').masks[1][1], raw[1][1])
    cr = pm.crop(np.array([1, 1, 4, 3]))
    assert (cr.height, cr.width) == (2, 3) and np.allclose(cr.masks[0][0][:2], [0, 0])
    assert pm.pad((8, 8)).height == 8
    assert np.allclose(pm.areas, [6.5, 1.5])
    car = pm.crop_and_resize(np.array([[1., 1, 3, 3]], np.float32), (4, 4), [0])
    assert np.allclose(car.masks[0][0][:4], [0, 0, 4, 0])
    assert len(PM([], 5, 5).crop_and_resize(np.zeros((0, 4), np.float32), (4, 4), [])) == 0
    assert PM([], 5, 5).to_ndarray().shape == (0, 5, 5)
    if not torch.cuda.is_available():
        with pytest.raises(NotImplementedError):
            pm.to_ndarray()
        with pytest.raises(NotImplementedError):
            pm.crop_and_resize_device(np.zeros((1, 4), np.float32), [(4, 4)], np.zeros(1, np.int64), 'cpu')


def test_next_row_ops_refuse_cpu_tensors():
    import dynamask_b200 as dm
    with pytest.raises(NotImplementedError):
        dm.SimpleRoIAlign(7, 0.25)(torch.zeros(1, 2, 8, 8), torch.zeros(1, 5))
    with pytest.raises(NotImplementedError):
        dm.refine_stage_instance_preds([torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 8, 8)])
    layer = dm.SimpleRoIAlign(14, 1.0 / 4)
    assert layer.output_size == (14, 14) and layer.spatial_scale == 0.25 and layer.aligned


def test_rle_compress_batch_host_equals_per_instance_strings():
    rng = np.random.default_rng(6)
    h, w = 23, 31
    masks = [(rng.random((h, w)) < p).astype(np.uint8) for p in (0.0, 1.0, 0.5, 0.1, 0.9)]
    trans = [_transitions(m) for m in masks]
    offs = np.zeros(len(masks) + 1, np.int64)
    np.cumsum([len(t) for t in trans], out=offs[1:])
    flat = np.ascontiguousarray(np.concatenate(trans), np.int32)
    lib = _lib.load()
    cap = 6 * flat.size + 8 * len(masks) + 8
    buf = np.empty(cap, np.uint8)
    so = np.empty(len(masks) + 1, np.int64)
    n = lib.dm_rle_compress_batch_host(ctypes.c_void_p(flat.ctypes.data), ctypes.c_void_p(offs.ctypes.data),
                                       len(masks), h * w, ctypes.c_void_p(buf.ctypes.data), cap,
                                       ctypes.c_void_p(so.ctypes.data))
    assert n == so[-1] and n > 0
    raw = buf[:n].tobytes()
    for i, t in enumerate(trans):
        assert raw[so[i]:so[i + 1]] == _compress(t, h * w)
    assert lib.dm_rle_compress_batch_host(ctypes.c_void_p(flat.ctypes.data), ctypes.c_void_p(offs.ctypes.data),
                                          len(masks), h * w, ctypes.c_void_p(buf.ctypes.data), 3,
                                          ctypes.c_void_p(so.ctypes.data)) == -1


def test_bench_helpers_degrade_without_a_gpu():
    """bench.py's host helpers must not fail (or change the process) where there is no NVML / GPU:
    the NUMA binding returns None and leaves the affinity alone, the clock sampler reports why it has
    no samples."""
    import bench
    aff = os.sched_getaffinity(0)
    got = bench.bind_to_gpu_numa_node(0)
    if got is None:
        assert os.sched_getaffinity(0) == aff
    else:                                   # a GPU box: bound to a non-empty subset
        assert 0 < got <= len(aff)
        os.sched_setaffinity(0, aff)
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(out) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
