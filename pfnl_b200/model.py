"""Drop-in host side of the PFNL forward path: same class/method names, argument meaning,
directory conventions and tensor layout as the reference's model/pfnl.py, with every
tensor op executed by libpfnl_b200.so on a B200 (PyTorch tensors are only the buffer
currency: device memory, streams, pinned host memory).

    reference                                   here
    PFNL.forward(x)               pfnl.py:39    PFNL.forward(x)        x [N,7,H,W,3] -> [N,1,4H,4W,3]
    sess.run(SR_test, feed_dict)  pfnl.py:252   PFNL.forward(numpy)    host in -> host out
    eval_mse                      pfnl.py:90    PFNL.eval_mse(sr, hr)  -> [N,1]
    test_video_truth              pfnl.py:203   same signature
    test_video_lr                 pfnl.py:264   same signature (alias: testvideo, README.md:31)
    testvideos                    pfnl.py:322   same signature
    eval()                        pfnl.py:94    same loop (eval_dir list, border 8, centre 15 step 32)
    load / save                   base_model.py:223-243   TF V2 checkpoints read/written without TF
    AVG_PSNR                      utils.py:216  PFNL.avg_psnr;  matlab/SSIM.m, compute_psnr.m -> PFNL.ssim_y / psnr_y
"""
from __future__ import annotations

import atexit
import ctypes as C
import glob
import os
import sys
import time
import weakref
from os.path import join

import numpy as np
import torch

from . import _lib, tf_checkpoint as _tfck, weights as _weights
from ._lib import check, lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_LIVE_ENGINES = weakref.WeakSet()


@atexit.register
def _close_live_engines():
    """Destroy every handle while the CUDA runtime is still up.  A handle finalised later - during
    interpreter shutdown, e.g. one kept alive by the traceback of an uncaught exception - would call
    into a runtime that is being torn down (observed: the process then hangs instead of exiting)."""
    for e in list(_LIVE_ENGINES):
        try:
            e.close()
        except Exception:
            pass


class Engine:
    """One libpfnl_b200 handle (one device).  Thin, typed wrapper over the C ABI."""

    def __init__(self, weights, device=0, precision="fp32", graphs=True):
        if not torch.cuda.is_available():
            raise RuntimeError("pfnl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.precision = precision
        prec = _lib.PRECISIONS[precision]
        st, keep = _weights.to_struct(weights)
        h = C.c_void_p()
        check(lib.pfnl_create(C.byref(h), self.device.index, C.byref(st), prec))
        del keep
        self._h = h
        _LIVE_ENGINES.add(self)
        check(lib.pfnl_set_graphs(self._h, 1 if graphs else 0))

    def close(self):
        if getattr(self, "_h", None):
            lib.pfnl_destroy(self._h)
            self._h = None

    def __del__(self):
        if sys.is_finalizing():  # too late for CUDA calls; the atexit hook has already closed live handles
            return
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -------------------------------------------------------------------------
    def _chk_in(self, t, ndim=None):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("expected a contiguous float32 CUDA tensor")
        if t.device != self.device:
            raise ValueError(f"tensor on {t.device}, engine on {self.device}")
        if ndim is not None and t.dim() != ndim:
            raise ValueError(f"expected {ndim} dims, got {tuple(t.shape)}")

    def _chk_out(self, t, shape):
        self._chk_in(t)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"out must have shape {tuple(shape)}, got {tuple(t.shape)}")

    @property
    def launches(self):
        return int(lib.pfnl_launch_count(self._h))

    PROF_KINDS = ("pack_tokens", "nonlocal", "conv0", "conv1_3x3", "conv10_1x1", "conv2_3x3", "convmerge1",
                  "tail", "other", "pfrb_flow", "nl_tc_kernel")

    def set_flow(self, enable):
        """Tensor-core precisions: PFRB stack as one persistent dataflow kernel (default) or two launches per block."""
        check(lib.pfnl_set_flow(self._h, 1 if enable else 0))

    def profile(self, enable):
        """Bracket every launch class with CUDA events (bypasses CUDA graphs while on)."""
        check(lib.pfnl_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """-> {class: (total_ms, launches)} since the last read (synchronises the device)."""
        ms = (C.c_double * len(self.PROF_KINDS))()
        cnt = (C.c_longlong * len(self.PROF_KINDS))()
        check(lib.pfnl_profile_read(self._h, ms, cnt))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.PROF_KINDS)}

    def reserve(self, n, h, w):
        check(lib.pfnl_reserve(self._h, n, h, w))

    def workspace_bytes(self, n, h, w):
        return int(lib.pfnl_workspace_bytes(_lib.PRECISIONS[self.precision], n, h, w))

    # -- whole forward -------------------------------------------------------------------
    def forward(self, x, out=None):
        self._chk_in(x, 5)
        n, f, h, w, c = x.shape
        if f != _lib.NUM_FRAMES or c != 3:
            raise ValueError(f"input must be [N,7,H,W,3], got {tuple(x.shape)}")
        if out is None:
            out = torch.empty((n, 1, h * 4, w * 4, 3), dtype=torch.float32, device=self.device)
        else:
            self._chk_out(out, (n, 1, h * 4, w * 4, 3))
        check(lib.pfnl_forward(self._h, _ptr(x), n, h, w, _ptr(out), _stream_ptr(self.device)))
        return out

    def forward_host(self, x, out=None):
        """numpy / CPU-tensor in, CPU tensor out (H2D + forward + D2H inside the call)."""
        ticket, out = self.forward_host_submit(x, out)
        self.forward_host_wait(ticket)
        return out

    def forward_host_submit(self, x, out=None):
        """Pipelined form of forward_host: starts H2D -> forward -> D2H for one batch and returns (ticket, out);
        `out` is complete after forward_host_wait(ticket).  Submitting the next batch before waiting overlaps the
        copies with compute (two staging slots: at most two batches in flight).  float64 input (what the
        reference feeds, pfnl.py:209,252) is narrowed inside the library on its way into pinned staging; the
        caller must keep `x` unchanged only until this call returns, `out` alive until the wait."""
        xt = torch.as_tensor(x)
        if xt.is_cuda:
            raise ValueError("forward_host takes host memory; use forward() for CUDA tensors")
        if xt.dtype not in (torch.float32, torch.float64):
            xt = xt.to(torch.float32)
        xt = xt.contiguous()
        if xt.dim() != 5:
            raise ValueError(f"input must be [N,7,H,W,3], got {tuple(xt.shape)}")
        n, f, h, w, c = xt.shape
        if f != _lib.NUM_FRAMES or c != 3:
            raise ValueError(f"input must be [N,7,H,W,3], got {tuple(xt.shape)}")
        if out is None:
            out = torch.empty((n, 1, h * 4, w * 4, 3), dtype=torch.float32)
        elif not (isinstance(out, torch.Tensor) and not out.is_cuda and out.dtype == torch.float32
                  and out.is_contiguous() and tuple(out.shape) == (n, 1, h * 4, w * 4, 3)):
            raise ValueError(f"out must be a contiguous float32 CPU tensor of shape {(n, 1, h * 4, w * 4, 3)}")
        ticket = C.c_int(-1)
        with torch.cuda.device(self.device):
            check(lib.pfnl_forward_host_submit(self._h, C.c_void_p(xt.data_ptr()), 1 if xt.dtype == torch.float64 else 0,
                                               n, h, w, C.c_void_p(out.data_ptr()), _stream_ptr(self.device),
                                               C.byref(ticket)))
        return ticket.value, out

    def forward_host_wait(self, ticket):
        with torch.cuda.device(self.device):
            check(lib.pfnl_forward_host_wait(self._h, ticket))

    def graph_stats(self):
        """(executables instantiated, executables re-pointed in place, executables cached)."""
        return tuple(int(lib.pfnl_graph_stats(self._h, k)) for k in range(3))

    def mse(self, sr, hr):
        self._chk_in(sr, 5)
        self._chk_in(hr, 5)
        # the reference's (SR - H)**2 (pfnl.py:90) would broadcast or raise on a shape mismatch; the kernel reads
        # both tensors with sr's extent, so anything but equal [N,1,4H,4W,3] shapes is rejected here
        if sr.shape != hr.shape or sr.shape[1] != 1 or sr.shape[-1] != 3:
            raise ValueError(f"sr and hr must both be [N,1,4H,4W,3]; got {tuple(sr.shape)} and {tuple(hr.shape)}")
        n, _, h4, w4, _ = sr.shape
        out = torch.empty((n,), dtype=torch.float32, device=self.device)
        check(lib.pfnl_mse(self._h, _ptr(sr), _ptr(hr), n, h4, w4, _ptr(out), _stream_ptr(self.device)))
        return out

    # -- stage-level ---------------------------------------------------------------------
    def pack_tokens(self, x):
        self._chk_in(x, 5)
        n, _, h, w, _ = x.shape
        out = torch.empty((n, (h // 2) * (w // 2), _lib.NL_CH), dtype=torch.float32, device=self.device)
        check(lib.pfnl_pack_tokens(self._h, _ptr(x), n, h, w, _ptr(out), _stream_ptr(self.device)))
        return out

    def nonlocal_block(self, tokens):
        self._chk_in(tokens, 3)
        n, l, c = tokens.shape
        if c != _lib.NL_CH:
            raise ValueError("tokens must be [N,L,84]")
        out = torch.empty_like(tokens)
        check(lib.pfnl_nonlocal(self._h, _ptr(tokens), n, l, _ptr(out), _stream_ptr(self.device)))
        return out

    def depth_to_space(self, x, block):
        self._chk_in(x, 4)
        n, h, w, c = x.shape
        out = torch.empty((n, h * block, w * block, max(c // (block * block), 1)), dtype=torch.float32,
                          device=self.device)
        check(lib.pfnl_depth_to_space(self._h, _ptr(x), n, h, w, c, block, _ptr(out), _stream_ptr(self.device)))
        return out

    def space_to_depth(self, x, block):
        self._chk_in(x, 4)
        n, h, w, c = x.shape
        out = torch.empty((n, max(h // block, 1), max(w // block, 1), c * block * block), dtype=torch.float32,
                          device=self.device)
        check(lib.pfnl_space_to_depth(self._h, _ptr(x), n, h, w, c, block, _ptr(out), _stream_ptr(self.device)))
        return out

    def conv2d(self, x, kernel, bias, act=False, residual=None):
        self._chk_in(x, 4)
        self._chk_in(kernel, 4)
        self._chk_in(bias, 1)
        n, h, w, ci = x.shape
        kh, kw, kci, co = kernel.shape
        if kh != kw or kci != ci:
            raise ValueError("kernel must be HWIO [k,k,Cin,Cout] matching the input channels")
        if residual is not None:
            self._chk_in(residual, 4)
        out = torch.empty((n, h, w, co), dtype=torch.float32, device=self.device)
        check(lib.pfnl_conv2d_nhwc(self._h, _ptr(x), n, h, w, ci, _ptr(kernel), _ptr(bias), kh, co, 1 if act else 0,
                                   _ptr(residual), _ptr(out), _stream_ptr(self.device)))
        return out

    def bicubic4(self, x):
        self._chk_in(x, 4)
        n, h, w, c = x.shape
        out = torch.empty((n, 4 * h, 4 * w, c), dtype=torch.float32, device=self.device)
        check(lib.pfnl_bicubic4(self._h, _ptr(x), n, h, w, c, _ptr(out), _stream_ptr(self.device)))
        return out

    # -- steps around the hot path (test_video_*), on the device ---------------------------
    def downsample4(self, hr, blur=None):
        """DownSample_4D (utils.py:169-192): hr [F,H,W,3] -> [F,ceil(H/4),ceil(W/4),3]."""
        self._chk_in(hr, 4)
        f, hh, ww, c = hr.shape
        if c != 3:
            raise ValueError("hr must be [F,H,W,3]")
        if blur is None:
            blur = gkern(13, 1.6)
        b = np.ascontiguousarray(np.asarray(blur, np.float32).reshape(169))
        out = torch.empty((f, (hh - 1) // 4 + 1, (ww - 1) // 4 + 1, 3), dtype=torch.float32, device=self.device)
        check(lib.pfnl_downsample4(self._h, _ptr(hr), f, hh, ww, C.c_void_p(b.ctypes.data), _ptr(out),
                                   _stream_ptr(self.device)))
        torch.cuda.current_stream(self.device).synchronize()  # `b` is host memory owned by this call
        return out

    def gather_windows(self, frames, first, count):
        """frames [F,h,w,3] -> clips [count,7,h,w,3] with edge clamping (pfnl.py:236-242)."""
        self._chk_in(frames, 4)
        f, hh, ww, c = frames.shape
        if c != 3:
            raise ValueError("frames must be [F,h,w,3]")
        out = torch.empty((count, _lib.NUM_FRAMES, hh, ww, 3), dtype=torch.float32, device=self.device)
        check(lib.pfnl_gather_windows(self._h, _ptr(frames), f, hh, ww, first, count, _ptr(out),
                                      _stream_ptr(self.device)))
        return out

    def quantize_u8(self, x):
        """round(clip(x*255,0,255)).astype(uint8) (pfnl.py:255-257)."""
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise ValueError("expected a contiguous float32 CUDA tensor")
        out = torch.empty(x.shape, dtype=torch.uint8, device=self.device)
        check(lib.pfnl_quantize_u8(self._h, _ptr(x), x.numel(), _ptr(out), _stream_ptr(self.device)))
        return out

    # -- evaluation metrics on the luma channel ------------------------------
    def msy(self, a, b, vmin=0.0, vmax=1.0, sp_border=8, round_y=False):
        """Per-frame mean squared Y difference (utils.py AVG_PSNR / matlab/compute_psnr.m):
        a, b [F,H,W,3] float32 CUDA -> [F] float64 CUDA."""
        self._chk_in(a, 4)
        self._chk_in(b, 4)
        if a.shape != b.shape or a.shape[-1] != 3:
            raise ValueError("expected two [F,H,W,3] tensors of the same shape")
        f, hh, ww, _ = a.shape
        out = torch.empty((f,), dtype=torch.float64, device=self.device)
        check(lib.pfnl_msy(self._h, _ptr(a), _ptr(b), f, hh, ww, vmin, vmax, sp_border, 1 if round_y else 0,
                           _ptr(out), _stream_ptr(self.device)))
        return out

    def ssim_y(self, a, b, vmin=0.0, vmax=1.0):
        """Per-frame mean SSIM of the luma as matlab/SSIM.m computes it: a, b [F,H,W,3] -> [F] float64."""
        self._chk_in(a, 4)
        self._chk_in(b, 4)
        if a.shape != b.shape or a.shape[-1] != 3:
            raise ValueError("expected two [F,H,W,3] tensors of the same shape")
        f, hh, ww, _ = a.shape
        out = torch.empty((f,), dtype=torch.float64, device=self.device)
        check(lib.pfnl_ssim_y(self._h, _ptr(a), _ptr(b), f, hh, ww, vmin, vmax, _ptr(out),
                              _stream_ptr(self.device)))
        return out

    def pfrb(self, blk, frames, n, h, w):
        """frames [N*7,H,W,64] -> one Progressive Fusion Residual Block (pfnl.py:66-71)."""
        self._chk_in(frames, 4)
        if tuple(frames.shape) != (n * _lib.NUM_FRAMES, h, w, 64):
            raise ValueError(f"frames must be [{n * _lib.NUM_FRAMES},{h},{w},64], got {tuple(frames.shape)}")
        out = torch.empty_like(frames)
        check(lib.pfnl_pfrb(self._h, blk, _ptr(frames), n, h, w, _ptr(out), _stream_ptr(self.device)))
        return out

    def conv0(self, inp21):
        """inp21 [N,H,W,21] (frame t = channels 3t..3t+2) -> conv0 of every frame [N*7,H,W,64] (pfnl.py:48,61-62)."""
        self._chk_in(inp21, 4)
        n, h, w, c = inp21.shape
        if c != 3 * _lib.NUM_FRAMES:
            raise ValueError("inp21 must be [N,H,W,21]")
        out = torch.empty((n * _lib.NUM_FRAMES, h, w, 64), dtype=torch.float32, device=self.device)
        check(lib.pfnl_conv0(self._h, _ptr(inp21), n, h, w, _ptr(out), _stream_ptr(self.device)))
        return out

    def convmerge1(self, frames, n, h, w):
        """frames [N*7,H,W,64] -> convmerge1 over the concat of the 7 frames [N,H,W,48] (pfnl.py:52,73-74)."""
        self._chk_in(frames, 4)
        if tuple(frames.shape) != (n * _lib.NUM_FRAMES, h, w, 64):
            raise ValueError(f"frames must be [{n * _lib.NUM_FRAMES},{h},{w},64], got {tuple(frames.shape)}")
        out = torch.empty((n, h, w, 48), dtype=torch.float32, device=self.device)
        check(lib.pfnl_convmerge1(self._h, _ptr(frames), n, h, w, _ptr(out), _stream_ptr(self.device)))
        return out


# -- image IO with the reference's conventions (utils.py:362-372) ---------------------------
def cv2_imread(path):
    import cv2
    img = cv2.imread(path)
    if img is None:
        raise IOError(f"cannot read {path}")
    if img.ndim == 3:
        img = img[:, :, ::-1]  # BGR -> RGB
    return img


def cv2_imsave(path, img):
    import cv2
    if img.ndim == 3:
        img = img[:, :, ::-1]  # RGB -> BGR
    return cv2.imwrite(path, np.ascontiguousarray(img))


def automkdir(path):
    os.makedirs(path, exist_ok=True)


def gkern(kernlen=13, nsig=1.6):
    """utils.py:95-102: 13x13 sigma=1.6 Gaussian (scipy's gaussian_filter of a dirac)."""
    from scipy.ndimage import gaussian_filter
    inp = np.zeros((kernlen, kernlen))
    inp[kernlen // 2, kernlen // 2] = 1
    return gaussian_filter(inp, nsig)


def downsample_4d(imgs, scale=4):
    """DownSample_4D (utils.py:169-192): REFLECT pad 6, depthwise 13x13 Gaussian, stride `scale`.
    Host-side for now (input preparation, the step before the hot path)."""
    blur = gkern(13, 1.6).astype(np.float32)
    x = np.asarray(imgs, np.float32)
    xp = np.pad(x, ((0, 0), (6, 6), (6, 6), (0, 0)), mode="reflect")
    n, hp, wp, c = xp.shape
    ho, wo = (hp - 13) // scale + 1, (wp - 13) // scale + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, (13, 13), axis=(1, 2))[:, ::scale, ::scale]
    win = win[:, :ho, :wo]
    return np.einsum("nhwcij,ij->nhwc", win, blur, optimize=True).astype(np.float32)


class PFNL:
    """Drop-in for the reference's `PFNL(VSR)` class on the inference path."""

    def __init__(self, weights=None, device=None, precision="fp32", graphs=True):
        # hyper-parameters kept with the reference's names and values (model/pfnl.py:21-37)
        self.num_frames = 7
        self.scale = 4
        self.in_size = 32
        self.gt_size = self.in_size * self.scale
        self.eval_in_size = [128, 240]
        self.batch_size = 16
        self.eval_basz = 4
        self.save_dir = './checkpoint/pfnl'
        self.log_dir = './pfnl.txt'
        self.eval_dir = './data/filelist_val.txt'   # model/pfnl.py:30
        self.global_step = 0                        # restored from the checkpoint when it holds one
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.precision = precision
        self.device_pipeline = True   # window gather / uint8 quantise / DownSample_4D on the device
        self._device = device
        self._graphs = graphs
        self._weights = None
        self._engine = None
        if weights is not None:
            self.load_weights(weights)

    # -- weights ---------------------------------------------------------------------------
    def load_weights(self, weights):
        """dict {tf_variable_name: array} or a path to an .npz of the same."""
        if isinstance(weights, (str, os.PathLike)):
            weights = _weights.load_npz(weights)
        self._weights = _weights.validate(weights)
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        return True

    def load(self, checkpoint_dir=None, step=None):
        """base_model.py:231-243: restore from `save_dir`; returns False (and keeps the initialiser
        values) when nothing is found.  Reads, in this order, the TensorFlow V2 checkpoint named by
        `<dir>/checkpoint` (what `tf.train.Saver.restore` reads; parsed without TF by tf_checkpoint.py -
        optimizer slots and other non-model variables are ignored) or `<dir>/pfnl.npz`."""
        d = checkpoint_dir or self.save_dir
        print(" [*] Reading SR checkpoints...")
        prefix = _tfck.latest_checkpoint(d) if os.path.isdir(d) else None
        if prefix is not None:
            names = list(_weights.variable_shapes())
            have = _tfck.list_variables(prefix)
            # the reference's step counter is an unnamed tf.Variable(initial_value=0) (pfnl.py:152): it sits in its
            # checkpoints under the key 'Variable' (int32); 'global_step' is what older saves of this class wrote
            step_key = next((k for k in ("Variable", "global_step") if k in have), None)
            tensors = _tfck.read_bundle(prefix, names + ([step_key] if step_key else []))
            if step_key is not None:
                self.global_step = int(np.asarray(tensors.pop(step_key)).reshape(-1)[0])
            else:  # base_model.py:227 names the files VSR-<step>
                tail = os.path.basename(prefix).rsplit("-", 1)[-1]
                if tail.isdigit():
                    self.global_step = int(tail)
            self.load_weights(tensors)
            print(" [*] Reading checkpoints...{} Success".format(os.path.basename(prefix)))
            return True
        p = join(d, "pfnl.npz")
        if os.path.exists(p):
            self.load_weights(p)
            print(" [*] Reading checkpoints...{} Success".format(os.path.basename(p)))
            return True
        print(" [*] Reading checkpoints... ERROR")
        return False

    def save(self, checkpoint_dir=None, step=None):
        """base_model.py:223-229: `<dir>/VSR-<step>.{index,data-00000-of-00001}` + `<dir>/checkpoint`, in
        TensorFlow's V2 format.  Holds the `nlvsr/*` model variables and the step counter as int32 under 'Variable'
        (the name the reference's unnamed counter gets, pfnl.py:152).  That is what the reference's inference-time
        Saver restores (test_video_* build a Saver over the model variables only); its train-time restore also
        expects the Adam slot variables, which an inference-only library does not have - not compatible with that."""
        d = checkpoint_dir or self.save_dir
        step = self.global_step if step is None else int(step)
        os.makedirs(d, exist_ok=True)
        if self._weights is None:
            self._weights = _weights.xavier_init()
        tensors = dict(self._weights)
        tensors["Variable"] = np.array(step, np.int32)
        name = "VSR-{}".format(step)
        _tfck.write_bundle(join(d, name), tensors)
        _tfck.write_checkpoint_state(d, name)
        return join(d, name)

    @property
    def engine(self):
        if self._engine is None:
            if self._weights is None:
                self._weights = _weights.xavier_init()
            self._engine = Engine(self._weights, self._device, self.precision, self._graphs)
        return self._engine

    # -- the hot path ----------------------------------------------------------------------
    def forward(self, x):
        """x [N,7,H,W,3] float32 in [0,1] -> [N,1,4H,4W,3] (model/pfnl.py:39-80).
        CUDA tensor in -> CUDA tensor out (asynchronous on the current stream);
        numpy / CPU tensor in -> same kind out (feed/fetch like sess.run, pfnl.py:252)."""
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return self.engine.forward(x.to(torch.float32).contiguous())
        out = self.engine.forward_host(x)
        return out if isinstance(x, torch.Tensor) else out.numpy()

    def eval_mse(self, sr, hr):
        """tf.reduce_mean((SR-H)**2, axis=[2,3,4]) (pfnl.py:90) -> [N,1]."""
        dev = self.engine.device
        srt = torch.as_tensor(sr, dtype=torch.float32).to(dev).contiguous()
        hrt = torch.as_tensor(hr, dtype=torch.float32).to(dev).contiguous()
        m = self.engine.mse(srt, hrt)[:, None]
        return m if isinstance(sr, torch.Tensor) and sr.is_cuda else m.cpu().numpy()

    @staticmethod
    def psnr(mse):
        """10*log10(1/mse) (pfnl.py:139)."""
        return 10.0 * np.log10(1.0 / np.asarray(mse, dtype=np.float64))

    # -- quality metrics on the luma channel -----------------------------------------------
    def _dev4(self, v):
        t = torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v)
        return t.to(device=self.engine.device, dtype=torch.float32).contiguous()

    def avg_psnr(self, vid_true, vid_pred, vmin=0, vmax=255, t_border=2, sp_border=8):
        """utils.py:216-246 (AVG_PSNR with RGB inputs): mean over the frames t_border..F-t_border of
        20*log10(255/rmse) of the luma difference, spatially cropped by sp_border."""
        m = self.engine.msy(self._dev4(vid_true), self._dev4(vid_pred), float(vmin), float(vmax), sp_border, False)
        m = m.cpu().numpy()
        m = m[t_border:m.shape[0] - t_border]
        return float(np.mean(20 * np.log10(255. / np.sqrt(m))))

    def psnr_y(self, img1, img2):
        """matlab/compute_psnr.m on uint8-range RGB images [H,W,3] or stacks [F,H,W,3] (values 0..255)."""
        a, b = self._dev4(img1), self._dev4(img2)
        if a.dim() == 3:
            a, b = a[None], b[None]
        m = self.engine.msy(a, b, 0.0, 255.0, 0, True).cpu().numpy()
        return 20 * np.log10(255.0 / np.sqrt(m))

    def ssim_y(self, img1, img2):
        """matlab/SSIM.m (default arguments) on uint8-range RGB images [H,W,3] or stacks [F,H,W,3]."""
        a, b = self._dev4(img1), self._dev4(img2)
        if a.dim() == 3:
            a, b = a[None], b[None]
        return self.engine.ssim_y(a, b, 0.0, 255.0).cpu().numpy()

    # -- validation loop (model/pfnl.py:94-149) -----------------------------------------------
    def eval(self, eval_dir=None):
        """The reference's `eval`: for every sequence listed in `eval_dir`, centre frames 15, 47, ... ->
        7-frame ground-truth windows cropped to eval_in_size*4 at border 8 -> LR by blur + x4 decimation
        (DownSample, utils.py:108-134 == DownSample_4D per frame) -> forward -> per-clip MSE -> PSNR; prints
        and appends the same log line.  Returns (psnr_avg, mse_avg) like the printed values."""
        print('Evaluating ...')
        if self._weights is None:
            self.load(self.save_dir)
        eng = self.engine
        border = 8
        in_h, in_w = self.eval_in_size
        out_h, out_w = in_h * self.scale, in_w * self.scale
        filenames = open(eval_dir or self.eval_dir, 'rt').read().splitlines()
        gt_list = [sorted(glob.glob(join(f, 'truth', '*.png'))) for f in filenames]
        center = 15
        batch_gt = []
        batch_cnt = 0
        mse_acc = None
        for gtlist in gt_list:
            max_frame = len(gtlist)
            for idx0 in range(center, max_frame, 32):
                index = np.array([i for i in range(idx0 - self.num_frames // 2, idx0 + self.num_frames // 2 + 1)])
                index = np.clip(index, 0, max_frame - 1).tolist()
                gt = [cv2_imread(gtlist[i]) for i in index]
                gt = [i[border:out_h + border, border:out_w + border, :].astype(np.float32) / 255.0 for i in gt]
                batch_gt.append(np.stack(gt, axis=0))
                if len(batch_gt) == self.eval_basz:
                    gt_dev = torch.as_tensor(np.stack(batch_gt, 0)).to(eng.device).contiguous()   # [B,7,oh,ow,3]
                    lr = eng.downsample4(gt_dev.view(-1, out_h, out_w, 3)).view(len(batch_gt), self.num_frames,
                                                                               in_h, in_w, 3)
                    sr = eng.forward(lr)
                    mid = self.num_frames // 2
                    mse_val = eng.mse(sr, gt_dev[:, mid:mid + 1].contiguous())[:, None].cpu().numpy()
                    mse_acc = mse_val if mse_acc is None else np.concatenate([mse_acc, mse_val], axis=0)
                    batch_gt = []
                    print('\tEval batch {} - {} ...'.format(batch_cnt, batch_cnt + self.eval_basz))
                    batch_cnt += self.eval_basz
        if mse_acc is None:
            raise ValueError("eval: fewer than eval_basz={} clips found under {}".format(self.eval_basz,
                                                                                        eval_dir or self.eval_dir))
        psnr_acc = 10 * np.log10(1.0 / mse_acc)
        mse_avg = np.mean(mse_acc, axis=0)
        psnr_avg = np.mean(psnr_acc, axis=0)
        print('Eval PSNR: {}, MSE: {}'.format(psnr_avg, mse_avg))
        with open(self.log_dir, 'a+') as f:
            mse_log = (mse_avg * 1e6).astype(np.int64) / (1e6)
            psnr_log = (psnr_avg * 1e6).astype(np.int64) / (1e6)
            f.write('{' + '"Iter": {} , "PSNR": {}, "MSE": {}'.format(self.global_step, psnr_log.tolist(),
                                                                     mse_log.tolist()) + '}\n')
        return psnr_avg, mse_avg

    # -- video harnesses (model/pfnl.py:203-332) -------------------------------------------
    def _window_list(self, lrs):
        max_frame = lrs.shape[0]
        lr_list = []
        for i in range(max_frame):
            index = np.array([i for i in range(i - self.num_frames // 2, i + self.num_frames // 2 + 1)])
            index = np.clip(index, 0, max_frame - 1).tolist()
            lr_list.append(np.array([lrs[j] for j in index]))
        return np.array(lr_list)

    def _run_video(self, lrs, save_path, part):
        """The loop of test_video_truth / test_video_lr (model/pfnl.py:236-262, 294-320).
        `lrs` [F,h,w,3]: numpy (host) or a float32 CUDA tensor.  With `device_pipeline` (default) the LR
        frames are uploaded ONCE, the 7-frame windows are gathered on the device, and the SR frames come
        back as uint8 (quantised on the device): 7x less H2D and 4x less D2H than feeding `lr_list`."""
        max_frame = lrs.shape[0]
        if max_frame == 0:
            return np.array([])
        if part > max_frame:
            part = max_frame
        num_once = max_frame // part if max_frame % part == 0 else max_frame // part + 1
        print('Save at {}'.format(save_path))
        print('{} Inputs With Shape {}'.format(lrs.shape[0], tuple(lrs.shape[1:])))
        eng = self.engine
        if self.device_pipeline:
            frames = torch.as_tensor(np.asarray(lrs) if not isinstance(lrs, torch.Tensor) else lrs,
                                     dtype=torch.float32).to(eng.device).contiguous()
        else:
            lr_list = self._window_list(np.asarray(lrs.cpu()) if isinstance(lrs, torch.Tensor) else lrs)
        all_time = []
        for i in range(part):
            first = i * num_once
            count = min(num_once, max_frame - first)
            if count <= 0:
                break
            st_time = time.time()
            if self.device_pipeline:
                clips = eng.gather_windows(frames, first, count)
                imgs = eng.quantize_u8(eng.forward(clips)).cpu().numpy()[:, 0]
            else:
                sr = self.forward(np.ascontiguousarray(lr_list[first:first + count], dtype=np.float32))
                imgs = [np.round(np.clip(sr[j][0] * 255., 0, 255), 0).astype(np.uint8) for j in range(sr.shape[0])]
            all_time.append(time.time() - st_time)
            for j in range(count):
                cv2_imsave(join(save_path, '{:0>4}.png'.format(first + j)), imgs[j])
        all_time = np.array(all_time)
        if max_frame > 0:
            mean_t = np.mean(all_time[1:]) if all_time.size > 1 else float('nan')
            print('spent {} s in total and {} s in average'.format(np.sum(all_time), mean_t))
        return all_time

    def test_video_truth(self, path, name='result', reuse=False, part=50):
        save_path = join(path, name)
        automkdir(save_path)
        inp_path = join(path, 'truth')
        imgs = sorted(glob.glob(join(inp_path, '*.png')))
        imgs = np.array([cv2_imread(i) for i in imgs]) / 255.
        if not reuse and self._weights is None:
            self.load(self.save_dir)
        if not len(imgs):
            return self._run_video(np.zeros((0, 0, 0, 3), np.float32), save_path, part)
        if self.device_pipeline:   # blur + x4 decimation on the device (DownSample_4D, utils.py:169-192)
            hr = torch.as_tensor(imgs, dtype=torch.float32).to(self.engine.device).contiguous()
            lrs = self.engine.downsample4(hr)
        else:
            lrs = downsample_4d(imgs, self.scale)
        return self._run_video(lrs, save_path, part)

    def test_video_lr(self, path, name='result', reuse=False, part=50):
        save_path = join(path, name)
        automkdir(save_path)
        inp_path = join(path, 'blur{}'.format(self.scale))
        imgs = sorted(glob.glob(join(inp_path, '*.png')))
        lrs = np.array([cv2_imread(i) for i in imgs]) / 255.
        if not reuse and self._weights is None:
            self.load(self.save_dir)
        return self._run_video(lrs, save_path, part)

    testvideo = test_video_lr  # README.md:31 spelling used by the other models

    def testvideos(self, path='/dev/f/data/video/test2/udm10', start=0, name='pfnl'):
        kind = sorted(glob.glob(join(path, '*')))
        kind = [k for k in kind if os.path.isdir(k)]
        reuse = False
        for k in kind:
            idx = kind.index(k)
            if idx >= start:
                if idx > start:
                    reuse = True
                datapath = join(path, k)
                self.test_video_truth(datapath, name=name, reuse=reuse, part=1000)
