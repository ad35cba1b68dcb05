"""Stand-in for ``nvdiffrast.torch.texture`` - TEST INFRASTRUCTURE ONLY.

nvdiffrast is an un-vendored, un-pinned third-party CUDA library (named only in
the reference's README.md:14) and is not installable offline.  The reference
calls it at exactly one site (networks/gdb_nerf/bundle_sampler.py:355-359):
``texture(tex[N,H,W,C], uv[N,S,1,2], mip_level_bias=lod[N,S,1],
boundary_mode='clamp', max_mip_level=L)`` with filter_mode 'auto'.

This file restates the published behaviour of that call (nvdiffrast
``texture.cu``: calculateMipLevel / indexTextureLinear / TextureFwdKernel /
MipBuildKernel) so the unmodified reference can run on the CPU to produce golden
vectors.  PARITY UNPINNED: no reference test pins this boundary.

Rules implemented
  * mip chain: level k+1 = 0.25*((a00+a10)+(a01+a11)) over 2x2 blocks of level
    k, k < L; every halved dimension must be even.
  * filter 'auto' + mip_level_bias -> linear-mipmap-linear; with no uv_da the
    level of detail is the bias alone: lod = clamp(bias, 0, L);
    l0 = floor(lod); l1 = min(l0+1, L); f = lod - l0.
  * per level: u_t = u*w - 0.5, clamped to [0, w-1]; i0 = floor(u_t);
    i1 = i0 + (0 if clamped-at-edge else 1); frac = u_t - i0;
    bilerp = lerp(lerp(a00,a10,fu), lerp(a01,a11,fu), fv), lerp(a,b,t)=a+t*(b-a).
  * out = a                    if lod == 0
        = a + f*(b - a)        otherwise (a from l0, b from l1).
"""
import torch


def _build_mips(tex, max_level):
    levels = [tex]
    for _ in range(max_level):
        t = levels[-1]
        n, h, w, c = t.shape
        if h % 2 or w % 2:
            raise ValueError("mip dims must stay even: %dx%d" % (h, w))
        q = t.view(n, h // 2, 2, w // 2, 2, c)
        levels.append(0.25 * ((q[:, :, 0, :, 0] + q[:, :, 0, :, 1]) + (q[:, :, 1, :, 0] + q[:, :, 1, :, 1])))
    return levels


def _bilerp_level(t, uv):
    n, h, w, c = t.shape
    u = (uv[..., 0] * w - 0.5).clamp(0.0, w - 1.0)
    v = (uv[..., 1] * h - 0.5).clamp(0.0, h - 1.0)
    edge_u = (u == 0.0) | (u == w - 1.0)
    edge_v = (v == 0.0) | (v == h - 1.0)
    iu0 = u.floor().long()
    iv0 = v.floor().long()
    iu1 = iu0 + (~edge_u).long()
    iv1 = iv0 + (~edge_v).long()
    fu = (u - iu0).unsqueeze(-1)
    fv = (v - iv0).unsqueeze(-1)
    flat = t.reshape(n, h * w, c)

    def take(iv, iu):
        idx = (iv * w + iu).reshape(n, -1, 1).expand(-1, -1, c)
        return torch.gather(flat, 1, idx).view(*uv.shape[:-1], c)

    a00, a10, a01, a11 = take(iv0, iu0), take(iv0, iu1), take(iv1, iu0), take(iv1, iu1)
    top = a00 + fu * (a10 - a00)
    bot = a01 + fu * (a11 - a01)
    return top + fv * (bot - top)


def texture(tex, uv, uv_da=None, mip_level_bias=None, mip=None, filter_mode='auto',
            boundary_mode='wrap', max_mip_level=None):
    if uv_da is not None or mip is not None or mip_level_bias is None:
        raise NotImplementedError("stand-in covers the reference's single call shape only")
    if boundary_mode != 'clamp' or filter_mode != 'auto' or max_mip_level is None:
        raise NotImplementedError("stand-in covers the reference's single call shape only")
    levels = _build_mips(tex, int(max_mip_level))
    L = len(levels) - 1
    lod = mip_level_bias.clamp(0.0, float(L))
    l0 = lod.floor()
    l1 = (l0 + 1.0).clamp(max=float(L))
    frac = (lod - l0).unsqueeze(-1)
    per_level = [_bilerp_level(t, uv) for t in levels]
    a = torch.zeros_like(per_level[0])
    b = torch.zeros_like(per_level[0])
    for k, val in enumerate(per_level):
        a = torch.where((l0 == k).unsqueeze(-1), val, a)
        b = torch.where((l1 == k).unsqueeze(-1), val, b)
    return torch.where((lod > 0.0).unsqueeze(-1), a + frac * (b - a), a)
