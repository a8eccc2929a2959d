"""Device-agnostic torch restatement of the reference's hot-path modules.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl reference).
Nothing under depthinspace_b200/ imports this file.

Why it exists next to the plain-C oracle: the reference path is a composition of torch
ops, and two of its arithmetic details depend on which torch backend runs them
(tensor/scalar division and grid_sample's un-normalisation differ between the CPU and
CUDA kernels).  Running the SAME op sequence with torch on the GPU box gives the
reference's CUDA arithmetic without needing /root/reference there, and running it on
the host cores gives the multi-threaded CPU baseline the bench reports
(`cpu_baseline.kind = "port"`).  tests/test_oracle_pinning.py checks, in the build
container, that every function here reproduces the imported reference bit-for-bit on CPU.

Each function cites the reference lines it restates.
"""
import torch
import torch.nn.functional as F

LOSS_TYPES = ("mse", "sad", "census_mse", "census_sad")


def lcn(x, radius=5, eps=0.05):
    """model/networks.py:667-689 -> (normalised, std)."""
    k = 2 * radius + 1
    ones = torch.ones(1, 1, k, k, dtype=x.dtype, device=x.device)

    def box(t):
        return F.conv2d(F.pad(t, (radius,) * 4, mode="reflect"), ones)

    s1 = box(x)
    mean = s1 / k ** 2
    s2 = box(x ** 2)
    std = torch.sqrt(torch.clamp(s2 / k ** 2 - mean ** 2 + 1e-6, min=0)) + eps
    return (x - mean) / std, std


def photometric(es, ta, block_size, type="mse", eps=0.1):
    """model/ext_functions.py:156-183 (the in-tree definition of the ext kernel)."""
    type = type.lower()
    if type not in LOSS_TYPES:
        raise Exception("invalid loss type")
    r = block_size // 2
    n, c, h, w = es.shape

    def windows(t):
        cols = F.unfold(F.pad(t, (r, r, r, r), mode="replicate"), kernel_size=block_size)
        return cols.view(n, c, -1, h, w)

    we, wt = windows(es), windows(ta)
    if type == "mse":
        per_tap = (we - wt) ** 2
    elif type == "sad":
        per_tap = (we - wt).abs()
    else:
        de = we - es.unsqueeze(2)
        dt = wt - ta.unsqueeze(2)
        he = 0.5 * (1 + de / torch.sqrt(de * de + eps))
        ht = 0.5 * (1 + dt / torch.sqrt(dt * dt + eps))
        delta = he - ht
        per_tap = delta * delta if type == "census_mse" else delta.abs()
    return per_tap.view(n, -1, h, w).sum(dim=1, keepdim=True) / block_size ** 2


def _pixel_grid(h, w, device, dtype):
    v, u = torch.meshgrid(torch.arange(h, device=device, dtype=dtype),
                          torch.arange(w, device=device, dtype=dtype), indexing="ij")
    return u, v


def pattern_warp(disp, pattern):
    """model/networks.py:356-367: x = u - disp, y = v, normalise, grid_sample(border)."""
    n, _, h, w = disp.shape
    u, v = _pixel_grid(h, w, disp.device, disp.dtype)
    xs = u.reshape(1, -1) - disp.contiguous().view(n, -1)
    ys = v.reshape(1, -1).expand(n, -1)
    gx = 2 * (xs / (w - 1) - 0.5)
    gy = 2 * (ys / (h - 1) - 0.5)
    grid = torch.stack((gx, gy), dim=-1).view(n, h, w, 2)
    return F.grid_sample(pattern.expand(n, -1, -1, -1), grid, padding_mode="border", align_corners=True)


def pattern_loss(disp, im, std, pattern, loss_type="census_sad", loss_eps=0.5, block_size=9,
                 output_mean=True, chunk=None):
    """RectifiedPatternSimilarityLoss.tforward, model/networks.py:354-377.
    ``pattern`` is the module's stored [1,1,H,W] (mean over the 3 channels, :344).
    ``chunk`` evaluates the unfold-based window loss a few frames at a time (the reference
    would need k*k times the batch in memory); the ratio is still formed over the batch."""
    proj = pattern_warp(disp, pattern)
    mask = torch.ones_like(im)
    if std is not None:
        mask = mask * std
    n = disp.shape[0]
    step = chunk or n
    if output_mean:
        num = 0
        for i in range(0, n, step):
            d = photometric(proj[i:i + step].contiguous(), im[i:i + step].contiguous(), block_size, loss_type, loss_eps)
            num = num + (mask[i:i + step] * d).sum()
        return num / mask.sum(), proj
    diff = torch.cat([photometric(proj[i:i + step].contiguous(), im[i:i + step].contiguous(), block_size,
                                  loss_type, loss_eps) for i in range(0, n, step)])
    return diff, proj


_SOBEL5 = [[-5, -4, 0, 4, 5], [-8, -10, 0, 10, 8], [-10, -20, 0, 20, 10], [-8, -10, 0, 10, 8], [-5, -4, 0, 4, 5]]
_SOBEL3 = [[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]]


def sobel(x, ksize=5, norm=False):
    """model/networks.py:697-730 -> [N,2,H,W] (gx, gy); weights rounded to float32 first."""
    base, div = (_SOBEL5, 240.0) if ksize == 5 else (_SOBEL3, 8.0)
    kx = (torch.tensor(base, dtype=torch.float64) / div).float().to(device=x.device, dtype=x.dtype)
    r = ksize // 2
    xp = F.pad(x, (r, r, r, r), mode="replicate")
    gx = F.conv2d(xp, kx[None, None])
    gy = F.conv2d(xp, kx.t().contiguous()[None, None])
    if norm:
        return torch.sqrt(gx ** 2 + gy ** 2 + 1e-8)
    return torch.cat((gx, gy), dim=1)


def smooth_loss(disp, im):
    """DisparitySmoothLoss.tforward, model/networks.py:419-431."""
    return (sobel(disp) * torch.exp(-(255 * sobel(im)).abs())).abs().mean()


def flow_warp(x, flow):
    """model/multi_frame_networks.py:83-99 (zeros padding, align_corners=True)."""
    h, w = x.shape[-2:]
    u, v = _pixel_grid(h, w, x.device, flow.dtype)
    grid = flow.clone().permute(0, 2, 3, 1)
    grid[..., 0] += u
    grid[..., 1] += v
    grid[..., 0] = 2 * (grid[..., 0] / (w - 1) - 0.5)
    grid[..., 1] = 2 * (grid[..., 1] / (h - 1) - 0.5)
    return F.grid_sample(x, grid, padding_mode="zeros", align_corners=True)


def fb_mask(flow_fwd, flow_bwd_warped, a=0.01, b=0.5):
    """forward-backward consistency mask, model/multi_frame_networks.py:205-207."""
    mag = (flow_fwd ** 2).sum(1, keepdim=True) + (flow_bwd_warped ** 2).sum(1, keepdim=True)
    diff = ((flow_fwd + flow_bwd_warped) ** 2).sum(1, keepdim=True)
    return (diff < a * mag + b).float()


def single_frame_loss(disps, im_lcn, std, ambient, pattern, chunk=None, pseudo_gt=None):
    """Photometric + smoothness part of single_frame_worker.Worker.loss_forward
    (model/single_frame_worker.py:101-125, 152-155): list of weighted 0-dim terms.
    disps: list of [N,1,H,W] (scale s weighted 1/2^s); smoothness on scale 0, weight 0.4."""
    vals = []
    for s, d in enumerate(disps):
        v, _ = pattern_loss(d, im_lcn, std, pattern, chunk=chunk)
        vals.append(v / (2 ** s))
    vals.append(smooth_loss(disps[0], ambient) * 0.4)
    if pseudo_gt is not None:
        for s, d in enumerate(disps):
            vals.append((d - pseudo_gt).abs().mean() * 0.1 / (2 ** s))
    return vals


def sgm_warmup_term(o, sgm_disp, noise):
    """The warm-up term of the first epochs on real data (model/single_frame_worker.py:158-163), with the reference's
    `1.5 * torch.randn(o.size()).cuda()` passed in as `noise`:  sum(|o - sgm + noise| * valid) / sum(valid) * 0.1."""
    valid_mask = (sgm_disp > 30).to(o.dtype)
    return torch.sum(torch.abs(o - sgm_disp + noise) * valid_mask) / torch.sum(valid_mask) * 0.1


def multi_frame_loss(disp, im_lcn, std, ambient, pattern, chunk=None, primary_disp=None):
    """Photometric + smoothness part of multi_frame_worker.Worker.loss_forward
    (model/multi_frame_worker.py:103-126, 160-165): smoothness weight 0.8."""
    v, _ = pattern_loss(disp, im_lcn, std, pattern, chunk=chunk)
    vals = [v, smooth_loss(disp, ambient) * 0.8]
    if primary_disp is not None:
        vals.append((disp - primary_disp).abs().mean() * 0.1)
    return vals


class FlowConsistency:
    """Restatement of ProjectionBaseLoss + Single/Multi_Frame_Flow_Consistency_Loss
    (model/networks.py:433-493, 554-661); K, Ki: [3,3] tensors as the workers pass them."""

    def __init__(self, K, Ki, im_height, im_width, clamp=-1, multi_frame=False):
        import numpy as np
        self.K = K.view(-1, 3, 3)
        self.h, self.w, self.clamp, self.multi_frame = im_height, im_width, clamp, multi_frame
        u, v = np.meshgrid(range(im_width), range(im_height))
        uv = np.stack((u, v, np.ones_like(u)), axis=2).reshape(-1, 3)
        self.ray = torch.from_numpy((uv @ Ki.numpy().T).reshape(1, -1, 3).astype(np.float32))
        self.u = torch.from_numpy(u.astype("float32"))
        self.v = torch.from_numpy(v.astype("float32"))

    def _to(self, ref):
        self.ray, self.K = self.ray.to(ref.device, ref.dtype), self.K.to(ref.device, ref.dtype)
        self.u, self.v = self.u.to(ref.device, ref.dtype), self.v.to(ref.device, ref.dtype)

    def project(self, depth, Ra, ta, Rb, tb):
        bs = depth.shape[0]
        xyz = depth.reshape(bs, -1, 1) * self.ray
        xyz = torch.bmm(xyz - ta.reshape(bs, 1, 3), Ra)
        xyz = torch.bmm(xyz, Rb.transpose(1, 2)) + tb.reshape(bs, 1, 3)
        uv = torch.bmm(xyz, self.K.transpose(1, 2).expand(bs, -1, -1))
        d = uv[:, :, 2:3]
        return uv[:, :, :2] / (F.relu(d) + 1e-12), d

    def direction(self, depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1, primary_depth1=None):
        self._to(depth0)
        _, d1 = self.project(depth0, R0, t0, R1, t1)
        d1 = d1.view(-1, 1, self.h, self.w)
        grid = flow0.permute(0, 2, 3, 1).clone()
        grid[..., 0] += self.u
        grid[..., 1] += self.v
        grid[..., 0] = 2 * (grid[..., 0] / (self.w - 1) - 0.5)
        grid[..., 1] = 2 * (grid[..., 1] / (self.h - 1) - 0.5)
        sample = lambda t: F.grid_sample(t, grid.detach() if not t.requires_grad else grid, padding_mode="zeros", align_corners=True)
        diff = (d1 - sample(depth1)).abs()
        orig = None
        if not self.multi_frame:
            if self.clamp > 0:
                diff = torch.clamp(diff, 0, self.clamp)
            orig = (diff.detach() < self.clamp).float()
        with torch.no_grad():
            f10 = sample(flow1.detach())
            fb = ((flow0 + f10) ** 2).sum(dim=1) < 0.5 + 0.02 * ((flow0 ** 2).sum(dim=1) + (f10 ** 2).sum(dim=1))
            mask = fb.to(depth0.dtype).unsqueeze(1)
            mask = mask * ((amb0 - sample(amb1.detach())).abs().mean(dim=1, keepdim=True) < 0.01).to(depth0.dtype)
            if self.multi_frame:
                uv0, _ = self.project(primary_depth1.detach(), R1, t1, R0, t0)
                uv0 = uv0.view(-1, self.h, self.w, 2).permute(0, 3, 1, 2)
                self_uv = torch.stack([self.u, self.v], dim=0).unsqueeze(0)
                mask = mask * (((sample(uv0) - self_uv) ** 2).sum(dim=1, keepdim=True) < 1).to(depth0.dtype)
        return (diff * mask).sum() / (mask.sum() + 1e-8), mask, orig

    def __call__(self, depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1, primary_depth0=None, primary_depth1=None):
        l0, m0, orig = self.direction(depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1, primary_depth1)
        l1, m1, _ = self.direction(depth1, depth0, R1, t1, R0, t0, flow1, flow0, amb1, amb0, primary_depth0)
        if self.multi_frame:
            return l0 + l1
        return l0 + l1, m0, m1, orig[0][0]


def disp_to_depth(disp, focal_length, baseline):
    """DispToDepth, model/networks.py:311-319."""
    return (baseline * focal_length) / (F.relu(disp) + 1e-12)


def conv3d_gather(xyz, feat, mask, ksize=3, stride=1, neighbors=9):
    """Conv3D.tforward up to the gathers, model/multi_frame_networks.py:469-501.
    -> (xyz_neighbors [M,nb,3], feat_neighbors [M,nb,C], neighbors_ind [M,nb,1])"""
    tl = xyz.shape[0]
    p = (ksize - 1) // 2
    pad = lambda t: F.pad(t, (p, p, p, p), mode="constant", value=0)
    unf = lambda t: pad(t).unfold(3, ksize, stride).unfold(4, ksize, stride).permute(1, 3, 4, 5, 6, 0, 2)
    xyz_u, feat_u, mask_u = unf(xyz), unf(feat), unf(mask)
    flat = lambda t: t.reshape(-1, ksize * ksize * tl, t.shape[-1])
    xyz_u, feat_u, mask_u = flat(xyz_u), flat(feat_u), flat(mask_u)
    plane = xyz_u / (xyz_u[..., 2:] + 1e-12)
    tidx = ((ksize ** 2) // 2) * tl
    xyz_local = xyz_u - xyz_u[:, tidx:tidx + 1, :]
    plane_local = plane - plane[:, tidx:tidx + 1, :]
    sq = (plane_local ** 2).sum(dim=-1, keepdim=True)
    key = (mask_u * sq) + (1 - mask_u) * (sq.max() + 1)
    _, ind = torch.topk(key, neighbors, dim=1, largest=False, sorted=False)
    xyz_nb = torch.gather(xyz_local, dim=1, index=ind.expand(-1, -1, xyz_local.shape[-1]))
    feat_nb = torch.gather(feat_u, dim=1, index=ind.expand(-1, -1, feat_u.shape[-1]))
    return xyz_nb, feat_nb, ind
