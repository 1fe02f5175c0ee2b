"""Out-of-bounds WRITE checks without compute-sanitizer (closed on this pool): every output the library
writes is carved out of a larger device buffer with guard bands on both sides, filled with a sentinel;
after the call the bands must be untouched and the payload must equal what the plugin surface returns
for the same inputs.  Calls go through the raw C ABI (``include/dynamask_sm100.h``) so that the test
owns every allocation.  Covers the TMA forward path (7 / 14 / 28), the cp.async ring (56), the X-first
and strip-walk backward, the fused and the two-launch paste, and the one-call paste -> RLE pipeline
with exactly-sized workspace / header / string buffers."""
import ctypes

import pytest
import torch

import synth

pytestmark = pytest.mark.gpu

STRIDES = [4, 8, 16, 32]
GUARD = 1 << 16          # bytes on each side
vp = ctypes.c_void_p


def dm():
    import dynamask_b200
    return dynamask_b200


def gen(seed):
    return torch.Generator().manual_seed(seed)


class Guarded:
    """``nbytes`` of payload (16-byte aligned) between two sentinel bands."""

    def __init__(self, nbytes, fill=0xA5):
        self.nbytes = int(nbytes)
        pad = (-self.nbytes) % 16
        self.buf = torch.full((GUARD + self.nbytes + pad + GUARD, ), fill, dtype=torch.uint8, device='cuda')
        self.fill = fill

    def view(self, dtype, shape=None):
        itemsize = torch.empty((), dtype=dtype).element_size()
        t = self.buf[GUARD:GUARD + self.nbytes].view(dtype)
        assert t.numel() * itemsize == self.nbytes
        return t if shape is None else t.view(shape)

    def ptr(self):
        return self.buf.data_ptr() + GUARD

    def check(self, what):
        torch.cuda.synchronize()
        lo = self.buf[:GUARD]
        hi = self.buf[GUARD + self.nbytes:]
        assert bool((lo == self.fill).all()), '%s: bytes BEFORE the buffer were written' % what
        assert bool((hi == self.fill).all()), '%s: bytes BEHIND the buffer were written' % what


def _level_args(tensors):
    ptrs = (vp * len(tensors))(*[t.data_ptr() for t in tensors])
    shapes = (ctypes.c_int32 * (4 * len(tensors)))(*[int(v) for t in tensors for v in t.shape])
    strides = (ctypes.c_int64 * (4 * len(tensors)))(*[int(v) for t in tensors for v in t.stride()])
    return ptrs, shapes, strides


@pytest.mark.parametrize('out_size,channels', [(7, 24), (14, 40), (28, 16), (56, 12)])
def test_roi_align_forward_and_backward_stay_inside_their_buffers(out_size, channels):
    from dynamask_b200 import _lib, ops
    lib = _lib.load()
    g = gen(500 + out_size)
    feats = [f.cuda() for f in synth.make_features(2, channels, 480, 672, g)]
    rois = synth.make_rois(2, 37, 480, 672, g).cuda()          # odd count, both images
    lvl = ops.assign(rois, None, 4, 56.0, 1)[0]
    K, P = rois.size(0), out_size
    scales = (ctypes.c_float * 4)(*[1.0 / s for s in STRIDES])
    ohw = (ctypes.c_int32 * 2)(P, P)
    stream = vp(torch.cuda.current_stream().cuda_stream)
    fptrs, fshapes, fstrides = _level_args(feats)
    for scratch in (True, False):                               # dynamic and static schedule
        sched = torch.zeros(16, dtype=torch.int32, device='cuda') if scratch else None
        gout = Guarded(4 * K * channels * P * P)
        out = gout.view(torch.float32, (K, channels, P, P))
        optrs = (vp * 1)(out.data_ptr())
        ostr = (ctypes.c_int64 * 4)(*out.stride())
        rc = lib.dm_roi_align_fwd(fptrs, fshapes, fstrides, scales, 4, vp(rois.data_ptr()), K, vp(lvl.data_ptr()),
                                  None, None, 1, ohw, optrs, ostr, 0, 1,
                                  vp(sched.data_ptr()) if scratch else None, stream)
        assert rc == 0
        gout.check('dm_roi_align_fwd %dx%d' % (P, P))
        ref = ops.multilevel_roi_align(feats, rois, [(P, P)], [1.0 / s for s in STRIDES], lvl=lvl)[0]
        assert torch.allclose(out, ref, rtol=1e-6, atol=1e-6)
        # backward: every level's gradient map between guard bands
        go = torch.randn(K, channels, P, P, generator=g).cuda()
        guards = [Guarded(4 * f.numel()) for f in feats]
        grads = [gd.view(torch.float32, tuple(f.shape)) for gd, f in zip(guards, feats)]
        gptrs, gshapes, gstrides = _level_args(grads)
        goptrs = (vp * 1)(go.data_ptr())
        gostr = (ctypes.c_int64 * 4)(*go.stride())
        rc = lib.dm_roi_align_bwd(gptrs, gshapes, gstrides, scales, 4, vp(rois.data_ptr()), K, vp(lvl.data_ptr()),
                                  None, None, 1, ohw, goptrs, gostr, 0, 1, 1,
                                  vp(sched.data_ptr()) if scratch else None, stream)
        assert rc == 0
        for l, gd in enumerate(guards):
            gd.check('dm_roi_align_bwd %dx%d level %d' % (P, P, l))
        want = ops.roi_align_backward([go], rois, lvl, None, None, [int(v) for f in feats for v in f.shape],
                                      [False] * 4, [P, P], [1.0 / s for s in STRIDES], 0, True)
        for l in range(4):
            assert torch.allclose(grads[l], want[l], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('hw,mode', [((301, 417), 'bool'), ((120, 150), 'bool'), ((301, 417), 'f32')])
def test_paste_stays_inside_its_canvas(hw, mode):
    from dynamask_b200 import _lib, ops
    lib = _lib.load()
    H, W = hw
    g = gen(600 + H)
    n = 7
    logits = synth.make_mask_logits(n, 28, g).cuda()
    boxes = synth.make_boxes(n, H, W, g).cuda()
    boxes[0] = torch.tensor([-30.0, -20.0, 40.0, 50.0])         # hangs over the top-left corner
    boxes[1] = torch.tensor([W - 25.0, H - 30.0, W + 40.0, H + 60.0])   # ... and over the bottom-right one
    es, out_mode, dtype = (4, ops.PASTE_F32, torch.float32) if mode == 'f32' else (1, ops.PASTE_BOOL, torch.uint8)
    gd = Guarded(n * H * W * es)
    stream = vp(torch.cuda.current_stream().cuda_stream)
    rc = lib.dm_paste_masks(vp(logits.data_ptr()), logits.stride(0), logits.stride(1), None, n, 28, 28, 1,
                            vp(boxes.data_ptr()), H, W, 0, 0, W, H, 0.5, out_mode, vp(gd.ptr()), stream)
    assert rc == 0
    gd.check('dm_paste_masks %s %dx%d' % (mode, H, W))
    ref = ops.paste_masks(logits, boxes, None, H, W, [0, 0, W, H], True, 0.5, out_mode)
    got = gd.view(dtype, (n, H, W))
    assert torch.equal(got, ref.view(dtype))


def test_paste_rle_strings_stays_inside_exactly_sized_buffers():
    from dynamask_b200 import _lib, ops
    lib = _lib.load()
    H, W = 203, 317
    g = gen(700)
    n = 11
    boxes = synth.make_boxes(n, H, W, g).cuda()
    stream = vp(torch.cuda.current_stream().cuda_stream)
    lin = (torch.arange(28, dtype=torch.float32) + 0.5) / 28 * 2 - 1
    clean = (4.0 * (1.0 - (lin[None, :] ** 2 + lin[:, None] ** 2)))[None, None].repeat(n, 1, 1, 1).cuda()
    noisy = torch.randn(n, 1, 28, 28, generator=g).cuda() * 3
    for masks in (clean, noisy):
        want = ops._paste_rle_two_pass(masks, boxes, None, H, W, (0, 0, W, H), True, 0.5)
        for record in (1, 0):
            for cap in (1 << 16, 64):                          # roomy, and far too small (status = 1: nothing written)
                ws = Guarded(lib.dm_paste_rle_strings_workspace(n, W, H, cap))
                hdr = Guarded(8 * (2 + n + 1))
                out = Guarded(6 * cap + 8 * n + 8)
                rc = lib.dm_paste_rle_strings(vp(masks.data_ptr()), masks.stride(0), masks.stride(1), None, n, 28, 28,
                                              1, vp(boxes.data_ptr()), H, W, 0, 0, W, H, 0.5, record, vp(ws.ptr()),
                                              cap, vp(hdr.ptr()), vp(out.ptr()), stream)
                assert rc == 0
                for gd, name in ((ws, 'workspace'), (hdr, 'header'), (out, 'strings')):
                    gd.check('dm_paste_rle_strings %s (capacity %d, record %d)' % (name, cap, record))
                h = hdr.view(torch.int64).cpu()
                if cap == 64:
                    assert int(h[0]) & 1 == 1 and int(h[1]) > cap
                    continue
                assert int(h[0]) & 1 == 0
                so = h[2:].tolist()
                raw = out.view(torch.uint8)[:so[-1]].cpu().numpy().tobytes()
                assert [raw[so[i]:so[i + 1]] for i in range(n)] == [r['counts'] for r in want]
