"""NWHead / NWNet — the API of the reference's nwhead/nw.py on the B200 CUDA path.

Two execution paths, both through libnw_sm100 (no PyTorch or CPU fallback):
  * tensor-core path  (SupportBank.forward -> nw_forward_class_lse): shared 2-D support with more than
    25 rows and no gradient required — inference (`predict`) and any no-grad call.  torch.cdist itself
    switches to the expanded |q|^2+|s|^2-2q.s form above 25 rows.
  * direct fp32 path  (nw_direct_*): exact differences, differentiable, per-query 3-D supports —
    episodic training (`forward`) and tiny supports.
A differentiable call with a big batch against a big shared support runs forward AND backward on the tensor
cores (nwhead_b200/backward.py).
"""
import weakref

import torch
import torch.nn as nn

from . import _abi
from ._abi import KIND, check, load, ptr, stream_of
from .backward import NWTensorFunction, wants_tensor_path
from .bank import SupportBank
from .kernel import get_kernel
from .support import SupportSetEval, SupportSetTrain

MM_PATH_MIN_ROWS = 26  # torch.cdist uses the matmul form when a side has more than 25 rows
_HALF_DTYPES = (torch.float16, torch.bfloat16)


class _LabelGuard:
    """Label validation for the direct path without a host/device synchronisation.

    The kernels store 1 into a status flag when a label is outside [0, n_classes) (what F.one_hot rejects,
    reference nwhead/nw.py:276); such a label contributes to no class and is never used as an index, so nothing
    can go out of bounds on the device.  The flag lives in PINNED HOST memory that the GPU writes directly (UVA:
    the pinned pointer is valid on the device), so inspecting it costs one host load: it is polled on EVERY
    forward and backward call, and `check(block=True)` (NWHead.check_labels) drains the stream first for callers
    that want the reference's raise-at-the-call behaviour."""

    def __init__(self, device):
        self.device = device
        self.flag = torch.zeros((1,), dtype=torch.int32).pin_memory()
        self.flag_ptr = self.flag.data_ptr()
        self._view = self.flag.numpy()

    def check(self, block=False):
        if block:
            torch.cuda.current_stream(self.device).synchronize()
        if self._view[0]:
            self._view[0] = 0
            raise RuntimeError("Class values must be smaller than num_classes.")


_guards = {}


def label_guard(device) -> _LabelGuard:
    key = (device.type, device.index)
    if key not in _guards:
        _guards[key] = _LabelGuard(device)
    return _guards[key]


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class _NWDirectFunction(torch.autograd.Function):
    """Differentiable NWHead.forward (reference nwhead/nw.py:266-289) on the direct fp32 kernels;
    backward is the closed form of SURVEY.md B.2 (nw_direct_backward).

    The episodic step (B=8, N=10) is 23 us of GPU work and host-bound, so this wrapper is written for few
    Python-level operations: two allocations in forward (the result, and one scratch block holding the scores, the
    row log-sum-exp AND the backward's workspace), two in backward (the gradients), pointers passed as plain
    integers, no views, no extra library calls."""

    @staticmethod
    def forward(ctx, x, sx, sy, logit_scale, kind, n_classes):
        lib = load()
        dev = x.device
        if not (x.is_cuda and sx.device == dev and sy.device == dev):
            dev = _abi.require_cuda(x, sx, sy)  # raises: a CPU tensor, or tensors on different devices
        xq, sxd, syd = _c(x), _c(sx), _c(sy)  # (grad mode is off inside forward: no detach needed)
        b, d = xq.shape
        batched = sxd.dim() == 3
        n = sxd.shape[-2]
        pairs = b * n
        scale = float(logit_scale.detach().exp()) if logit_scale is not None else 1.0
        guard = label_guard(dev)
        guard.check()
        logp = torch.empty((b, n_classes), dtype=torch.float32, device=dev)
        # scratch: scores (b*n) | row_lse (b) | backward workspace (nw_direct_backward_workspace_elems)
        ws = 0
        if any(ctx.needs_input_grad):  # nw_direct_backward_workspace_elems; the episodic case needs no library call
            ws = (pairs + (pairs if batched else n) + b if n <= 1024 else
                  lib.nw_direct_backward_workspace_elems(b, d, n, int(batched)))
        aux = torch.empty((pairs + b + ws,), dtype=torch.float32, device=dev)
        p_aux = aux.data_ptr()
        check(lib.nw_direct_forward(KIND[kind], scale, xq.data_ptr(), b, d, sxd.data_ptr(), n, int(batched),
                                    syd.data_ptr(), int(syd.dim() == 2), n_classes, p_aux, logp.data_ptr(),
                                    p_aux + 4 * pairs, guard.flag_ptr, stream_of(dev)), "nw_direct_forward")
        ctx.save_for_backward(xq, sxd, syd, aux, logp)
        ctx.kind, ctx.n_classes, ctx.scale = kind, n_classes, scale
        ctx.has_scale = logit_scale is not None
        return logp

    @staticmethod
    def backward(ctx, grad_out):
        lib = load()
        xq, sxd, syd, aux, logp = ctx.saved_tensors
        dev = xq.device
        label_guard(dev).check()  # polled on every call: raises as soon as a finished forward has flagged a bad label
        b, d = xq.shape
        batched = sxd.dim() == 3
        n = sxd.shape[-2]
        pairs = b * n
        need_q, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_scale = ctx.has_scale and ctx.needs_input_grad[3]
        g = grad_out if (grad_out.dtype == torch.float32 and grad_out.is_contiguous()) else grad_out.float().contiguous()
        gq = torch.empty_like(xq) if (need_q or not need_s) else None
        gs = torch.empty_like(sxd) if need_s else None
        gscale = torch.empty((b,), dtype=torch.float32, device=dev) if need_scale else None
        p_aux = aux.data_ptr()
        check(lib.nw_direct_backward(KIND[ctx.kind], ctx.scale, xq.data_ptr(), b, d, sxd.data_ptr(), n, int(batched),
                                     syd.data_ptr(), int(syd.dim() == 2), ctx.n_classes, p_aux, p_aux + 4 * pairs,
                                     logp.data_ptr(), g.data_ptr(), p_aux + 4 * (pairs + b),
                                     None if gq is None else gq.data_ptr(), None if gs is None else gs.data_ptr(),
                                     None if gscale is None else gscale.data_ptr(), stream_of(dev)),
              "nw_direct_backward")
        return (gq if need_q else None, gs, None, gscale.sum() if need_scale else None, None, None)


class NWHead(nn.Module):
    BANK_CACHE_SIZE = 2  # banks built from raw (sx, sy) tensors are kept while the tensors stay unmodified

    def __init__(self, kernel, n_classes, precision="auto", backward_path="auto"):
        super().__init__()
        self.kernel = kernel
        self.n_classes = n_classes
        self.precision = precision
        # differentiable calls: 'direct' = exact-difference fp32 kernels, 'tensor' = tcgen05 forward + backward
        # (bf16 operands, nwhead_b200/backward.py), 'auto' = tensor for big batches against big shared supports
        if backward_path not in ("auto", "direct", "tensor"):
            raise ValueError(f"unknown backward_path {backward_path!r}")
        self.backward_path = backward_path
        self._bank_cache = []

    def _bank_for(self, sx, sy, kind, precision=None):
        """The reference hands the SAME support tensors to the head on every predict call (nwhead/nw.py:156-160).
        Building the device bank (sort check, centring, bf16 conversion) once per tensor version instead of once
        per call keeps that usage pattern fast; in-place modification bumps `_version` and invalidates the entry."""
        # identity of the live tensor OBJECTS (weak references), not their addresses: a freed support's memory is
        # routinely handed to the next one (knn mode builds a new support per batch)
        precision = precision or self.precision
        if sx.is_inference() or sy.is_inference():
            # inference-mode tensors have no version counter: nothing to validate a cached bank against
            return SupportBank.build(sx, sy, self.n_classes, kind, precision)
        key = (sx._version, sy._version, kind, precision, self.n_classes)
        for ref_x, ref_y, k, bank in self._bank_cache:
            if ref_x() is sx and ref_y() is sy and k == key:
                return bank
        bank = SupportBank.build(sx, sy, self.n_classes, kind, precision)
        self._bank_cache = [e for e in self._bank_cache if e[0]() is not None and e[1]() is not None]
        self._bank_cache.append((weakref.ref(sx), weakref.ref(sy), key, bank))
        del self._bank_cache[:-self.BANK_CACHE_SIZE]
        return bank

    @staticmethod
    def check_labels(device="cuda") -> None:
        """Waits for the direct-path work queued on `device` and raises RuntimeError if any of it saw a support
        label outside [0, n_classes) — the blocking form of the check every call polls without blocking."""
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        label_guard(dev).check(block=True)

    def _kind(self):
        kind = getattr(self.kernel, "kind", None)
        if kind not in KIND:
            raise NotImplementedError(f"kernel {type(self.kernel).__name__} has no CUDA implementation")
        return kind

    def forward(self, x, sx, sy=None):
        """
        Computes Nadaraya-Watson head given query x, support x, and support y tensors.
        :param x: Query data (b, feat_dim)
        :param sx: Support data (num_support, feat_dim) or (b, num_support, feat_dim), or a prebuilt
            SupportBank (then sy is ignored).
        :param sy: Support targets (num_support) or (b, num_support), int64
        :return: log of softmaxed probabilities (b, num_classes)
        """
        kind = self._kind()
        if x.dtype in _HALF_DTYPES:  # a featurizer under autocast: torch.cdist upcasts these too
            x = x.float()
        if isinstance(sx, SupportBank):
            # inference-only (the reference's eval step runs predict() with gradients disabled, train.py:408).
            # A differentiable call needs the fp32 support rows: hand NWHead the raw (sx, sy) tensors instead.
            if torch.is_grad_enabled() and x.requires_grad:
                raise RuntimeError("a SupportBank is inference-only; pass the raw (sx, sy) support tensors for a "
                                   "differentiable call (NWNet.predict does this automatically)")
            return sx.forward_auto(x, self.kernel.scale_value())
        _abi.require_cuda(x, sx, sy)
        if sx.dtype in _HALF_DTYPES:
            sx = sx.float()
        if x.dtype != torch.float32 or sx.dtype != torch.float32:
            raise TypeError("NWHead expects float32 features")  # the reference raises on fp64 too
        if sy.dtype != torch.int64:
            raise RuntimeError("one_hot is only applicable to index tensor of type LongTensor.")
        logit_scale = self.kernel.logit_scale if kind == "clip" else None
        needs_grad = torch.is_grad_enabled() and (
            x.requires_grad or sx.requires_grad or (logit_scale is not None and logit_scale.requires_grad))
        n = sx.shape[-2]
        if x.shape[0] == 0:  # empty query batch, as the reference: an empty (0, n_classes) result
            return x.new_empty((0, self.n_classes))
        if n == 0:
            raise ValueError("NWHead needs at least one support row")
        if needs_grad and wants_tensor_path(x.shape[0], n, sx.dim(), self.backward_path):
            # a support that is not being trained keeps its bank (and its transposed copy) across steps
            bank = None if sx.requires_grad else self._bank_for(sx, sy, kind, "bf16")
            return NWTensorFunction.apply(x, sx, sy, logit_scale, kind, self.n_classes, bank)
        if needs_grad or sx.dim() == 3 or n < MM_PATH_MIN_ROWS:
            if needs_grad and x.shape[-1] + self.n_classes > _abi.DIRECT_BACKWARD_MAX_D_PLUS_C:
                # fail before the forward, not at .backward()
                raise NotImplementedError(
                    f"the differentiable path supports feat_dim + n_classes <= {_abi.DIRECT_BACKWARD_MAX_D_PLUS_C}, "
                    f"got {x.shape[-1]} + {self.n_classes}")
            return _NWDirectFunction.apply(x, sx, sy, logit_scale, kind, self.n_classes)
        return self._bank_for(sx, sy, kind).forward(x, self.kernel.scale_value())


class NWNet(nn.Module):
    def __init__(self,
                 featurizer,
                 n_classes,
                 support_dataset=None,
                 feat_dim=None,
                 proj_dim=0,
                 kernel_type='euclidean',
                 train_type='random',
                 n_way=None,
                 n_shot=1,
                 n_shot_random=1,
                 n_shot_full=100,
                 n_shot_cluster=1,
                 n_neighbors=10,
                 env_array=None,
                 debug_mode=False,
                 device='cuda:0',
                 return_mask=False,
                 precision='auto',
                 ):
        '''
        Top level NW net class (constructor of the reference, nwhead/nw.py:12-105, plus `precision`).

        :param featurizer: Feature extractor
        :param n_classes: Number of classes in dataset
        :param support_dataset: Pytorch Dataset with a .targets attribute of categorical labels
        :param feat_dim: Output dimension of featurizer
        :param proj_dim: If > 0, adds a linear projection down to proj_dim after featurizer
        :param kernel_type: Type of kernel to use
        :param train_type: Type of training strategy, choose from ['random', 'irm']
        :param n_way: Number of classes to put in support during training
        :param n_shot: Number of datapoints per class to sample for support during training
        :param n_shot_random / n_shot_full / n_shot_cluster: per-class support sizes of the eval modes
        :param n_neighbors: Number of neighbors for knn inference
        :param device: Device used for computation (must be a CUDA device)
        :param return_mask: If true, also returns a mask telling if the query class is in the support
        :param precision: 'auto' | 'bf16' | 'bf16x3' operand precision of the tensor-core path
        '''
        super().__init__()
        self.featurizer = featurizer
        self.train_type = train_type
        self.n_way = n_way
        self.debug_mode = debug_mode
        self.n_classes = n_classes
        self.n_shot = n_shot
        self.n_shot_random = n_shot_random
        self.n_shot_full = n_shot_full
        self.n_shot_cluster = n_shot_cluster
        self.n_neighbors = n_neighbors
        self.env_array = env_array
        self.device = device
        self.return_mask = return_mask
        self.precision = precision
        if support_dataset is not None:
            assert hasattr(support_dataset, 'targets'), 'Support set must have .targets attribute'

        if proj_dim > 0:
            assert feat_dim is not None, 'Feature dimension must be specified'
            self.featurizer = nn.Sequential(self.featurizer, nn.Linear(feat_dim, proj_dim))

        self.kernel = get_kernel(kernel_type)
        self.nwhead = NWHead(kernel=self.kernel, n_classes=n_classes, precision=precision)

        if support_dataset is not None:
            self.support_train = SupportSetTrain(support_dataset, self.n_classes, self.train_type, self.n_shot,
                                                 n_way=self.n_way, env_array=self.env_array)
            self.process_support_eval(support_dataset)

    def process_support_eval(self, support_dataset):
        '''Processes support dataset into SupportSet object.'''
        self.support_eval = SupportSetEval(support_dataset, self.n_classes, self.n_shot_random,
                                           self.n_shot_full, n_shot_cluster=self.n_shot_cluster,
                                           n_neighbors=self.n_neighbors, env_array=self.env_array,
                                           kernel_type=self.kernel.kind, precision=self.precision)

    def precompute(self):
        '''Precomputes all support features, cluster centroids, and random iterator.
        Call before running inference.  The bank stays on the GPU (bf16 + norms + labels).'''
        assert not self.featurizer.training
        sinfo = self._compute_all_support_feats()
        self.full_feat = sinfo[0]
        self.full_y = sinfo[1]
        self.support_eval.build_infer_iters(*sinfo)

    def predict(self, x, mode='random'):
        '''
        Perform prediction given test images.

        :param x: Input datapoints (bs, nch, l, w)
        :param mode: Inference mode. One of ['random', 'full', 'cluster', 'ensemble', 'knn', 'hnsw']
        '''
        qfeat = self.featurizer(x)
        # the reference's predict is differentiable (nwhead/nw.py:127-160): with gradients flowing into the
        # query features the head gets the raw fp32 support rows (direct path), otherwise the device bank
        raw = torch.is_grad_enabled() and qfeat.requires_grad
        support = self.support_eval.get_support(mode, x=qfeat, raw=raw)
        if self.debug_mode:
            print('qx shape:', x.shape)
        if mode == 'ensemble':
            # mean over environments of the per-environment class probabilities (reference nwhead/nw.py:143-154)
            probs = 0
            for bank in support:
                probs = probs + (self.nwhead(qfeat, *bank) if isinstance(bank, tuple) else self.nwhead(qfeat, bank)).exp()
            out = torch.log(probs / len(support))
        elif isinstance(support, SupportBank):
            out = self.nwhead(qfeat, support)
        else:
            sfeat, sy = support
            out = self.nwhead(qfeat, sfeat.to(x.device), sy.to(x.device))
        if self.return_mask:
            return out, torch.full((len(x),), True)
        return out

    def forward(self, x, y, metadata=None, support_data=None):
        '''
        Forward pass using images for query and support (episodic training step).

        :param x: Input datapoints (bs, nch, l, w)
        :param y: Corresponding labels (bs)
        :param metadata: Corresponding metadata (bs)
        :param support_data: Optional (sx, sy, sm) tuple for functional implementation
        '''
        if support_data is not None:
            sx, sy, sm = support_data
        else:
            sx, sy, sm = self.support_train.get_support(y)
        if sm is None:
            sm = torch.zeros_like(sy)
        sx, sy, sm = sx.to(x.device), sy.to(x.device), sm.to(x.device)

        batch_size = len(x)
        feats = self.featurizer(torch.cat((x, sx), dim=0))
        qfeat, sfeat = feats[:batch_size], feats[batch_size:]
        isin = torch.isin(y, sy)
        if self.debug_mode:
            print('qx shape:', x.shape, 'sx shape:', sx.shape)
            print('qy:', y, 'sy:', sy, 'qy in sy:', isin)
        out = self.nwhead(qfeat, sfeat, sy)
        if self.return_mask:
            return out, isin
        return out

    def _compute_all_support_feats(self):
        """Runs the featurizer over the class-balanced support loaders (reference nwhead/nw.py:213-243);
        features stay on the device instead of being copied to the host per batch."""
        feats, labels, meta = [], [], []
        sep_feats, sep_labels, sep_meta = [], [], []
        for loader in self.support_eval.support_loaders:
            env_feats, env_labels, env_meta = [], [], []
            for qimg, qlabel, qmeta in loader:
                feat = self.featurizer(qimg.to(self.device)).detach().float()
                env_feats.append(feat)
                env_labels.append(qlabel.to(self.device))
                env_meta.append(torch.as_tensor(qmeta))
            feats += env_feats
            labels += env_labels
            meta += env_meta
            sep_feats.append(torch.cat(env_feats, dim=0))
            sep_labels.append(torch.cat(env_labels, dim=0))
            sep_meta.append(torch.cat(env_meta, dim=0))
        return (torch.cat(feats, dim=0), torch.cat(labels, dim=0), torch.cat(meta, dim=0), sep_feats, sep_labels,
                sep_meta)

    def get_neighbors(self, x, k=None, exact=True):
        '''Returns indices of nearest neighbors of x in the support set, nearest first
        (reference nwhead/nw.py:245-249: the full ranking; pass k to keep only the first k).
        exact=True  : fp32 exact-difference scores (nw_direct_scores) + full ranking, bit-exact with the
                      reference on ties-free data.  With k given and an euclidean bank of TOPK_EXACT_MIN_ROWS rows
                      or more, the same ranking comes from SupportBank.topk_exact (tensor-core search over blocks
                      of 64 rows, fp32 re-rank of the certified candidates) without the (B, N) score matrix.
        exact=False : tensor-core scores against the precomputed bank (SupportBank.topk) for large banks /
                      batches; ranking accuracy is that of the bank's operand precision.'''
        from .utils import rank_rows

        qfeat = self.featurizer(x).detach()
        bank = self.support_eval.full_bank
        if not exact:
            return bank.topk(qfeat, len(bank) if k is None else k, self.kernel.scale_value())
        if k is not None and bank.kind == "euclidean" and len(bank) >= self.TOPK_EXACT_MIN_ROWS and k <= 1024:
            # large bank: tensor-core candidate search + exact fp32 re-rank (same ranking, no (B, N) matrix)
            return bank.topk_exact(qfeat, k, self.full_feat)
        distances = self.kernel(qfeat, self.full_feat)
        return rank_rows(distances, k)

    TOPK_EXACT_MIN_ROWS = 65536
