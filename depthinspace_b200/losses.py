"""Loss assembly of the reference's workers, restricted to the hot path (photometric + smoothness
+ auxiliary L1 terms); mirrors

    single_frame_worker.Worker.loss_forward   reference model/single_frame_worker.py:101-165
    multi_frame_worker.Worker.loss_forward    reference model/multi_frame_worker.py:103-175

Same weights and the same list-of-0-dim-tensors return convention (the worker sums it,
model/worker.py:522).  The geometric (flow-consistency) terms are appended by the caller.
"""
import os

import torch

from . import _ops
from .networks import (DisparitySmoothLoss, DispToDepth, Multi_Frame_Flow_Consistency_Loss,
                       RectifiedPatternSimilarityLoss, Single_Frame_Flow_Consistency_Loss)


class _L1Mean(torch.autograd.Function):
    """torch.mean(torch.abs(o - target)) with the gradient w.r.t. o produced by the same pass."""

    @staticmethod
    def forward(ctx, o, target, group):
        out3, sgn = _ops.l1_forward(o, target, want_grad=ctx.needs_input_grad[0])
        if group is not None:      # mean over the WHOLE data-parallel batch: all-reduce (sum, count), then divide
            from .parallel import all_reduce_sum_
            nd = out3[:2].contiguous()
            all_reduce_sum_(nd, group)
            out3 = torch.cat((nd, (nd[0] / nd[1]).reshape(1)))
        ctx.save_for_backward(sgn, out3)
        return out3[2].clone()

    @staticmethod
    def backward(ctx, g):
        sgn, out3 = ctx.saved_tensors
        return _ops.scale_by_device_scalar(sgn, g, out3[1:2]).view_as(sgn), None, None


def l1_mean(o, target, process_group=None):
    """mean |o - target| (gradient to o only: the targets are data); with a process group the mean runs over all ranks."""
    return _L1Mean.apply(o.contiguous(), target.detach(), process_group)


class _MaskedL1Mean(torch.autograd.Function):
    """sum(|o - target + noise| * valid) / sum(valid), valid = (target > threshold): the SGM warm-up term
    (reference model/single_frame_worker.py:158-163), value and gradient w.r.t. o from one pass."""

    @staticmethod
    def forward(ctx, o, target, noise, threshold, group):
        out3, sgn = _ops.masked_l1_forward(o, target, noise, threshold, want_grad=ctx.needs_input_grad[0])
        if group is not None:
            from .parallel import all_reduce_sum_
            nd = out3[:2].contiguous()
            all_reduce_sum_(nd, group)
            out3 = torch.cat((nd, (nd[0] / nd[1]).reshape(1)))
        ctx.save_for_backward(sgn, out3)
        return out3[2].clone()

    @staticmethod
    def backward(ctx, g):
        sgn, out3 = ctx.saved_tensors
        return _ops.scale_by_device_scalar(sgn, g, out3[1:2]).view_as(sgn), None, None, None, None


def masked_l1_mean(o, target, noise=None, threshold=30.0, process_group=None):
    return _MaskedL1Mean.apply(o.contiguous(), target.detach(), None if noise is None else noise.detach(), float(threshold),
                               process_group)


def _merge(x):
    """[tl,bs,C,H,W] -> [tl*bs,C,H,W] (model/multi_frame_networks.py:36-37); 4-D tensors and None pass through."""
    if x is None:
        return None
    return x.contiguous().view(-1, *x.shape[-3:]) if x.dim() == 5 else x


class _GeometricTerms(torch.autograd.Function):
    """All flow-consistency terms of one track batch (the pair loop of the workers) as ONE autograd node on the
    disparity maps: forward = DispToDepth + two fused launches per frame pair, backward = one launch that scales and
    sums the 4 gradient planes of every pair into d/d disp of the frames they belong to, DispToDepth's derivative
    included.  Autograd would otherwise run, per step, 12 two-plane combines, the zero-fills / strided copies of the
    `depth[i]` indexing, 10 gradient accumulations and DispToDepth's backward."""

    @staticmethod
    def forward(ctx, disp_tl, primary_disp, R, t, amb, K, ray, clamp, multi_frame, bf, weight, group, *flows):
        ctx.set_materialize_grads(False)
        tl = disp_tl.shape[0]
        disp_tl = disp_tl.contiguous()
        depth = bf / (torch.relu(disp_tl) + 1e-12)                     # DispToDepth, reference model/networks.py:311-319
        primary = bf / (torch.relu(primary_disp) + 1e-12) if multi_frame else None
        need = ctx.needs_input_grad[0]
        planes, frame_of, nd = [], [], []
        k = 0
        for i in range(tl):
            for j in range(i + 1, tl):
                f_ij, f_ji = flows[2 * k], flows[2 * k + 1]
                k += 1
                a3, _, _, gA0, gA1 = _ops.flow_consistency_dir(depth[i], depth[j], R[i], t[i], R[j], t[j], f_ij, f_ji, amb[i],
                                                               amb[j], K, ray, clamp, primary[j] if multi_frame else None,
                                                               False, False, need, need)
                b3, _, _, gB1, gB0 = _ops.flow_consistency_dir(depth[j], depth[i], R[j], t[j], R[i], t[i], f_ji, f_ij, amb[j],
                                                               amb[i], K, ray, clamp, primary[i] if multi_frame else None,
                                                               False, False, need, need)
                nd += [a3[:2], b3[:2]]
                if need:
                    planes += [gA0, gA1, gB1, gB0]
                    frame_of += [i, j, j, i]
        nd = torch.stack(nd) if nd else disp_tl.new_zeros((0, 2))          # [2 * pairs, (num, den)]
        if group is not None and k:   # the ratios run over the whole data-parallel batch: ONE all-reduce for every pair
            from .parallel import all_reduce_sum_
            all_reduce_sum_(nd, group)
        ratio = (nd[:, 0] / (nd[:, 1] + 1e-8)).view(-1, 2).sum(dim=1) * weight     # :599 / :653, times 0.2 / #pairs
        ctx.frame_of, ctx.bf, ctx.weight, ctx.n_pairs = frame_of, bf, weight, k
        dens = nd[:, 1].repeat_interleave(2) if need else disp_tl.new_zeros(0)     # (A, A, B, B) per pair
        ctx.save_for_backward(disp_tl, dens, *planes)
        return tuple(ratio.unbind(0))

    @staticmethod
    def backward(ctx, *g_vals):
        disp_tl, dens, *planes = ctx.saved_tensors
        if not planes:
            return (None,) * (12 + 2 * ctx.n_pairs)
        zero = dens.new_zeros(())
        g = torch.stack([zero if gv is None else gv.reshape(()) for gv in g_vals])       # [pairs]
        scale = (g.repeat_interleave(4) * ctx.weight / (dens + 1e-8)).contiguous()        # one factor per gradient plane
        grad = _ops.geometric_grad_combine(planes, ctx.frame_of, scale, disp_tl, ctx.bf)
        return (grad,) + (None,) * (11 + 2 * ctx.n_pairs)


class _HotPathLoss(torch.nn.Module):
    smooth_weight = None

    ge_class = None

    def __init__(self, im_height, im_width, pattern, loss_type='census_sad', loss_eps=0.5, block_size=9,
                 process_group=None, K=None, Ki=None, focal_length=None, baseline=None, ge_clamp=0.1):
        super().__init__()
        self.ph_loss = RectifiedPatternSimilarityLoss(im_height, im_width, pattern, loss_type, loss_eps,
                                                      block_size=block_size, return_pattern_proj=False,
                                                      process_group=process_group)
        self.disparity_loss = DisparitySmoothLoss(process_group=process_group)
        # geometric (flow-consistency) terms are optional: they need the intrinsics and the stereo baseline
        self.process_group = process_group
        self.ge_loss = (self.ge_class(K, Ki, im_height, im_width, clamp=ge_clamp, process_group=process_group)
                        if K is not None else None)
        self.d2d = DispToDepth(float(focal_length), float(baseline)) if focal_length is not None else None
        self._side_streams = {}

    # ---- smoothness on a side stream ------------------------------------------------------------------------------------
    # The photometric kernel's 2048 CTAs take 4.6 waves of the 444 resident slots: two fifths of its last wave are idle
    # SM time.  The smoothness kernel does not depend on it, so it is launched right behind it on a second stream and its
    # CTAs fill that tail.  DIS_OVERLAP_SMOOTH=0 (or overlap_smoothness = False) keeps everything on one stream.
    overlap_smoothness = os.environ.get("DIS_OVERLAP_SMOOTH", "1") != "0"

    def _fork_point(self, disp0, amb):
        """Event on the current stream BEFORE the photometric launch: what the side stream has to wait for.  Both inputs of
        the smoothness term must exist by then, in the layout the kernel reads (no copy may be queued behind the event)."""
        # (small batches are launch-bound: the extra event / stream calls would cost more than the overlap returns)
        if not (self.overlap_smoothness and amb.is_cuda and disp0.numel() >= (1 << 22) and disp0.is_contiguous() and amb.is_contiguous()
                and disp0.dtype == torch.float32 and amb.dtype == torch.float32) or torch.cuda.is_current_stream_capturing():
            return None
        bufs = _ops.smooth_loss_buffers(disp0, disp0.requires_grad and torch.is_grad_enabled())   # before the event, see _SmoothLoss
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(amb.device))
        return ev, bufs

    def _smooth_term(self, disp0, amb, fork):
        if fork is None:
            return self.disparity_loss(disp0, amb) * self.smooth_weight
        dev = disp0.device
        side = self._side_streams.get(dev)
        if side is None:
            side = self._side_streams[dev] = torch.cuda.Stream(device=dev)
        return self.disparity_loss(disp0, amb, fork, side) * self.smooth_weight

    def _geometric_terms(self, disp_tl, R, t, amb, flow_out, primary_disp=None):
        """The pair loop of the workers (single_frame_worker.py:127-149, multi_frame_worker.py:128-157):
        every unordered frame pair of a track, weight 0.2 / (tl (tl-1) / 2)."""
        tl = disp_tl.shape[0]
        ge_num = tl * (tl - 1) / 2
        multi_frame = primary_disp is not None
        K, ray = self.ge_loss._consts(disp_tl)
        flows = []
        for i in range(tl):
            for j in range(i + 1, tl):
                flows += [flow_out[f'flow_{i}{j}'].detach(), flow_out[f'flow_{j}{i}'].detach()]
        clamp = -1.0 if multi_frame else self.ge_loss.clamp        # the multi-frame variant never clamps (:564-601)
        return list(_GeometricTerms.apply(disp_tl, primary_disp.detach() if multi_frame else None, R, t, amb.detach(), K, ray,
                                          float(clamp), multi_frame, self.d2d.baseline_focal_length, 0.2 / ge_num,
                                          self.process_group, *flows))


    # ---- fused value + gradient (no autograd graph, no scaling passes) -------------------------------------------------
    def _value_and_grad(self, out, im_lcn, std, ambient, l1_target, l1_weights, global_frames=None):
        """Loss terms and their FINAL gradients w.r.t. every disparity map in one pass per kernel.
        The assembly knows its own weights (1/2^s, smooth_weight, ...), so the kernels store weight * d term / d disp
        directly: the denominators are known before the launch (sum of sigma: one streaming reduction of `std`; the
        smoothness count is 2*N*H*W), which removes autograd's five gradient-scaling passes over the batch.
        -> (list of 0-dim terms exactly like forward(), list of gradients shaped like `out`).  Continue into the network
        with  torch.autograd.backward(out, grads).  Geometric terms are not part of this path (use forward()).
        global_frames: frames of the whole data-parallel batch (all ranks); if omitted with a process group it is
        all-reduced and read back (one host sync).  Nothing else touches the host: the call can be graph-captured."""
        from .parallel import all_reduce_sum_
        if not isinstance(out, (tuple, list)):
            out = [out]
        ph = self.ph_loss
        type_id = _ops.loss_type_id(ph.loss_type)
        S = len(out)
        if S not in (1, 2, 4):
            raise ValueError("value_and_grad covers 1, 2 or 4 scales; use forward() otherwise")
        im, std_m, amb = _merge(im_lcn)[:, 0:1].contiguous(), _merge(std), _merge(ambient)
        disps = [_merge(o).detach() for o in out]
        dev, group = im.device, ph.process_group
        ph.pattern = ph.pattern.to(device=dev, dtype=torch.float32)
        n_local = disps[0].shape[0]
        hw = disps[0].shape[-1] * disps[0].shape[-2]
        if group is None:
            n_global = n_local
        elif global_frames is not None:
            n_global = int(global_frames)
        else:
            count = torch.full((1,), float(n_local), device=dev)
            all_reduce_sum_(count, group)
            n_global = int(round(float(count)))
        if std_m is not None:
            den = _ops.abs_sum(std_m)[0:1].clone()
            if group is not None:
                all_reduce_sum_(den, group)
        else:
            den = torch.full((1,), float(n_global * hw), device=dev)
        key = (S, str(dev))
        if getattr(self, "_w_ph_key", None) != key:      # constant weights 1 / 2^s, uploaded once
            self._w_ph, self._w_ph_key = torch.tensor([1.0 / 2 ** s for s in range(S)], device=dev), key
        w_ph = self._w_ph
        scale = (w_ph / den).contiguous()
        if S == 1:
            o3, _, _, g0 = _ops.pattern_loss_forward(disps[0], im, std_m, ph.pattern, ph.block_size, type_id, ph.loss_eps,
                                                     False, False, True, grad_scale=scale)
            out3, grads = o3.unsqueeze(0), [g0]
        else:
            out3, grads = _ops.pattern_loss_multi_forward(disps, im, std_m, ph.pattern, ph.block_size, type_id, ph.loss_eps, True,
                                                          grad_scale=scale)
        # smoothness on scale 0: mean over 2 * N * H * W elements (:118-124)
        per_frame = 2.0 * hw
        s3, _ = _ops.smooth_loss_forward(disps[0], amb.contiguous(), True, grad_scale=self.smooth_weight / (per_frame * n_global),
                                         accumulate_into=grads[0])
        sums = [out3[:, 0], s3[0:1]]                                   # S photometric numerators, the smoothness sum
        l1_terms = []
        if l1_target is not None:                                      # pseudo-GT / primary-disparity L1 terms
            tgt = _merge(l1_target).detach()
            for s, wgt in l1_weights:
                o3, sgn = _ops.l1_forward(disps[s], tgt, True)
                cnt = float(disps[s].numel() // n_local * n_global)    # elements of the whole batch: no reduction needed
                sums.append(o3[0:1])
                l1_terms.append(wgt / cnt)
                grads[s].add_(sgn.view_as(grads[s]), alpha=wgt / cnt)
        packed = torch.cat(sums)
        if group is not None:          # every numerator of the step in ONE all-reduced vector (S + 1 + #L1 floats)
            all_reduce_sum_(packed, group)
        vals = list((packed[:S] / den * w_ph).unbind(0))              # reference :108-115, weights 1 / 2^s
        vals.append(packed[S] * (self.smooth_weight / (per_frame * n_global)))
        for i, f in enumerate(l1_terms):
            vals.append(packed[S + 1 + i] * f)
        return vals, [g.view_as(o) for g, o in zip(grads, out)]


class SingleFrameLoss(_HotPathLoss):
    """out: list of per-scale disparities (all full resolution, model/networks.py:290-295).
    With R, t ([tl,bs,3,3], [tl,bs,3]) and flow_out ({'flow_ij': [bs,2,H,W]}) the 6 x 2 geometric terms are added
    in the reference's position (after smoothness, before the pseudo-GT terms); `out` must then be [tl,bs,1,H,W]."""
    ge_class = Single_Frame_Flow_Consistency_Loss
    smooth_weight = 0.4

    def forward(self, out, im_lcn, std, ambient, pseudo_gt=None, R=None, t=None, flow_out=None, sgm_disp=None,
                sgm_noise=None):
        """sgm_disp: adds the warm-up term of the first epochs on real data (:158-163), one per scale, weight 0.1:
        sum(|o - sgm + noise| * (sgm > 30)) / sum(sgm > 30).  sgm_noise: list of per-scale noise tensors; when omitted
        1.5 * randn is drawn on the device (the reference draws on the CPU generator: a different stream, same law)."""
        if not isinstance(out, (tuple, list)):
            out = [out]
        im, std, amb = _merge(im_lcn)[:, 0:1], _merge(std), _merge(ambient)
        disps = [_merge(o) for o in out]
        fork = self._fork_point(disps[0], amb)
        ph = self.ph_loss.forward_multi(disps, im, std)                      # :108-115, all scales fused
        vals = [v / (2 ** s) for s, v in enumerate(ph)]
        vals.append(self._smooth_term(disps[0], amb, fork))                  # :118-124 (scale 0 only)
        if flow_out is not None:                                      # :127-149
            vals += self._geometric_terms(out[0], R, t, ambient, flow_out)
        if pseudo_gt is not None:                                     # :152-155 (DIS-FTSF)
            for s, o in enumerate(out):
                vals.append(l1_mean(o, pseudo_gt, self.process_group) * 0.1 / (2 ** s))
        if sgm_disp is not None:                                      # :158-163 (warm-up, real data)
            for s, o in enumerate(out):
                noise = sgm_noise[s] if sgm_noise is not None else 1.5 * torch.randn_like(o)
                vals.append(masked_l1_mean(o, sgm_disp, noise, 30.0, self.ph_loss.process_group) * 0.1)
        return vals

    def value_and_grad(self, out, im_lcn, std, ambient, pseudo_gt=None, global_frames=None):
        """Same terms as forward() (without geometric terms) plus d(sum of terms)/d out[s], see _value_and_grad."""
        n = len(out) if isinstance(out, (tuple, list)) else 1
        return self._value_and_grad(out, im_lcn, std, ambient, pseudo_gt, [(s, 0.1 / 2 ** s) for s in range(n)],
                                    global_frames)


class MultiFrameLoss(_HotPathLoss):
    ge_class = Multi_Frame_Flow_Consistency_Loss
    smooth_weight = 0.8

    def forward(self, out, im_lcn, std, ambient, primary_disp=None, R=None, t=None, flow_out=None, warmup=True,
                sgm_disp=None, sgm_noise=None):
        if not isinstance(out, (tuple, list)):
            out = [out]
        im, std, amb = _merge(im_lcn)[:, 0:1], _merge(std), _merge(ambient)
        disps = [_merge(o) for o in out]
        fork = self._fork_point(disps[0], amb)
        ph = self.ph_loss.forward_multi(disps, im, std)                      # :110-117
        vals = [v / (2 ** s) for s, v in enumerate(ph)]
        vals.append(self._smooth_term(disps[0], amb, fork))                  # :120-126
        if flow_out is not None:                                      # :128-157 (needs primary_disp)
            vals += self._geometric_terms(out[0], R, t, ambient, flow_out, primary_disp=primary_disp)
        if primary_disp is not None and warmup:                       # :160-165 (first two epochs)
            vals.append(l1_mean(out[0], primary_disp, self.process_group) * 0.1)
        if sgm_disp is not None:                                      # :167-173 (warm-up on real data, scale 0 only)
            noise = sgm_noise if sgm_noise is not None else 1.5 * torch.randn_like(out[0])
            vals.append(masked_l1_mean(out[0], sgm_disp, noise, 30.0, self.ph_loss.process_group) * 0.1)
        return vals

    def value_and_grad(self, out, im_lcn, std, ambient, primary_disp=None, warmup=True, global_frames=None):
        """Photometric + smoothness (+ primary-disparity L1) terms of forward() and d(sum)/d out, see _value_and_grad;
        the geometric terms stay on forward()."""
        tgt = primary_disp if (primary_disp is not None and warmup) else None
        return self._value_and_grad(out, im_lcn, std, ambient, tgt, [(0, 0.1)], global_frames)
