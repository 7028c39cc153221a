"""
CPU oracle of the benchmark log-density (TEST INFRASTRUCTURE ONLY; same rules as pm_oracle.py).

Restates, in torch-CPU float64 on top of pm_oracle, the chain that `montecosmo_b200/model.py:FieldModel` runs on the
GPU: model.py:640-679 (prior, precond='real' or 'fourier', bricks.py:290-320), bricks.py:152-157 (white2lin), bricks.py:358-362 (linear Lagrangian
bias, NGP read), nbody.py:634-667 / 967-1002 (lpt / nbody_bf), bricks.py:781-792 (flat-sky RSD in cell units),
model.py:802-809 (nufft paint, irfftn), and a Gaussian likelihood.  Autograd gives the reference gradient.
"""
import numpy as np
import torch

from . import pm_oracle as O


def evolve(white, transfer, cosmo, mesh_shape, evolution="nbody", n_steps=5, a_start=0.0, a_obs=1.0, lpt_order=2,
           paint_order=2, interlace_order=2, paint_deconv=True, paint_shape=None, b1=1.0, rsd=True,
           los=(0.0, 0.0, 1.0), precond="real", checkpoint=False, state_out=None):
    """`checkpoint`: recompute the BullFrog steps in the backward pass (memory only).  `state_out`: a dict that receives
    the particle state ahead of the paint (q, pos, vel, weights), for the full-size fixtures."""
    mesh_shape = tuple(mesh_shape)
    paint_shape = mesh_shape if paint_shape is None else tuple(paint_shape)
    # samp2base_mesh, bricks.py:300-309: sample in real space (rfftn) or in Fourier space (rg2cgh)
    dk = (torch.fft.rfftn(white) if precond == "real" else O.rg2cgh(white)) * O._t(transfer)
    q = O.regular_pos(mesh_shape)
    weights = 1.0
    if b1 != 0.0:
        delta = torch.fft.irfftn(dk, s=mesh_shape)
        weights = 1.0 + b1 * O.a2g(cosmo, a_obs) * O.read(q, delta, 1)
    if evolution == "lpt":
        dpos, vel = O.lpt(cosmo, dk, q, a_obs, lpt_order, 1)
        pos = q + dpos
    else:
        pos, vel = O.nbody_bf(cosmo, dk, q, a_start, a_obs, n_steps, paint_order, lpt_order, paint_deconv=False,
                              checkpoint=checkpoint)
        pos, vel = pos[-1], vel[-1]
    if state_out is not None:
        state_out.update(q=q, pos=pos.detach(), vel=vel.detach(),
                         weights=weights.detach() if isinstance(weights, torch.Tensor) else weights)
    if rsd:
        l = O._t(np.asarray(los))
        pos = pos + (vel * l).sum(-1, keepdim=True) * (O.a2g(cosmo, a_obs) * O.a2f(cosmo, a_obs)) * l
    gxy = O.nufft(pos, mesh_shape, paint_shape, weights, paint_order, interlace_order, paint_deconv=paint_deconv)
    if paint_shape != mesh_shape:
        gxy = O.chreshape(gxy, O.r2chshape(paint_shape))
    return torch.fft.irfftn(gxy, s=paint_shape)


def logpdf(white, obs, transfer, cosmo, mesh_shape, sigma_obs=1.0, **kw):
    gxy = evolve(white, transfer, cosmo, mesh_shape, **kw)
    return -0.5 * ((gxy - O._t(obs)) ** 2).sum() / sigma_obs**2 - 0.5 * (white**2).sum()


def value_and_force(white, obs, transfer, cosmo, mesh_shape, **kw):
    w = O._t(white).clone().requires_grad_(True)
    lp = logpdf(w, obs, transfer, cosmo, mesh_shape, **kw)
    (g,) = torch.autograd.grad(lp, w)
    return lp.detach(), g


RH = 2997.92458  # jax_cosmo.constants.rh, h^-1 Mpc (bricks.py:125)
_PNG_KEYS = ("fNL_bp", "fNL_bpd", "fNL_bpd2", "fNL_bps2", "fNL_bn2p")


def trans_phi2delta(cosmo, kmesh, kpow, a=1.0):
    """bricks.py:108-127 with a tabulated (k, P) normalised to sigma8 = 1 (the Eisenstein-Hu branch of lin_power lives in
    jax_cosmo and is outside the path): transfer from the primordial potential to the linear density, on kmesh."""
    ks, pows = (np.asarray(x, dtype=np.float64) for x in kpow)
    pow_lin = pows * float(cosmo.sigma8) ** 2
    pow_large = ks ** float(cosmo.n_s)
    lin_trans = (pow_lin / pow_large / (pow_lin[0] / pow_large[0])) ** 0.5
    a_md = 1.0 / 11.0
    growth_md = float(O.a2g(cosmo, a_md)) / a_md
    trans = 2.0 * RH ** 2 * ks ** 2 * lin_trans * (float(O.a2g(cosmo, a)) / growth_md) / (3.0 * float(cosmo.Omega_m))
    km = np.asarray(kmesh, dtype=np.float64)
    return np.interp(km.reshape(-1), ks, trans, left=0.0, right=0.0).reshape(km.shape)


def _safe_div_t(x, y):
    y = O._t(y)
    return torch.where(y == 0, torch.zeros_like(x), x / torch.where(y == 0, torch.ones_like(y), y))


def add_png(cosmo, fNL, lin_mesh, box_size, kpow):
    """bricks.py:129-141."""
    lin_mesh = lin_mesh if isinstance(lin_mesh, torch.Tensor) else O._t(lin_mesh, O.C128)
    shape = O.ch2rshape(tuple(lin_mesh.shape))
    kmesh = sum(k ** 2 for k in O.rfftk(shape, box_size)) ** 0.5
    t = trans_phi2delta(cosmo, kmesh, kpow)
    phi = torch.fft.irfftn(_safe_div_t(lin_mesh, t), s=shape)
    phi2 = phi ** 2
    phi = phi + fNL * (phi2 - phi2.mean())
    return O._t(t) * torch.fft.rfftn(phi)


def lagrangian_bias(cosmo, pos, a, box_size, lin_mesh, bias, read_order=2, png=None, png_type=None, kpow=None):
    """bricks.py:327-452, float64: (weights, dvel), or (weights, dvel, phi) when png_type is not None (the primordial
    non-Gaussianity terms of bricks.py:411-438, with a tabulated kpow)."""
    b = {k: bias.get(k, 0.0) for k in ("b1", "b2", "bs2", "b3", "bds2", "bs3", "bn2", "bnpar")}
    lin_mesh = lin_mesh if isinstance(lin_mesh, torch.Tensor) else O._t(lin_mesh, O.C128)
    shape = O.ch2rshape(tuple(lin_mesh.shape))
    g = O.a2g(cosmo, a)
    g = g.reshape(-1) if isinstance(g, torch.Tensor) and g.numel() > 1 else g
    kvec = [O._t(k) for k in O.rfftk(shape, box_size)]
    kmesh2 = sum(k ** 2 for k in kvec)
    pot = lin_mesh * O._t(O.invlaplace_hat(O.rfftk(shape, box_size)))
    rd = lambda m: O.read(pos, m, read_order)
    irf = lambda m: torch.fft.irfftn(m, s=shape)
    delta_pos = rd(irf(lin_mesh)) * g
    w = 1.0 + b["b1"] * delta_pos
    d2 = delta_pos ** 2
    sigma2 = d2.mean()
    delta2_pos = d2 - sigma2
    w = w + b["b2"] * delta2_pos / 2
    sh = {}
    for i in range(2):
        nabi = 1j * kvec[i]
        sh[(i, i)] = irf(nabi ** 2 * pot - lin_mesh / 3)
        for j in range(i + 1, 3):
            sh[(i, j)] = irf(nabi * (1j * kvec[j]) * pot)
    sh[(2, 2)] = -(sh[(0, 0)] + sh[(1, 1)])
    a_, b_, c_ = sh[(0, 0)], sh[(1, 1)], sh[(2, 2)]
    d_, e_, f_ = sh[(0, 1)], sh[(0, 2)], sh[(1, 2)]
    shear2_pos = rd(a_ ** 2 + b_ ** 2 + c_ ** 2 + 2 * (d_ ** 2 + e_ ** 2 + f_ ** 2)) * g ** 2 - 2 / 3 * sigma2
    w = w + b["bs2"] * shear2_pos
    w = w + b["b3"] * (delta_pos ** 3 - 3 * sigma2 * delta_pos) / 6
    w = w + b["bds2"] * delta_pos * shear2_pos
    shear3 = 3 * (a_ * (b_ * c_ - f_ ** 2) - d_ * (d_ * c_ - e_ * f_) + e_ * (d_ * f_ - b_ * e_))
    w = w + b["bs3"] * rd(shear3) * g ** 3
    w = w + b["bn2"] * rd(irf(-kmesh2 * lin_mesh)) * g
    phi = None
    if png_type is not None:
        f = {k: (png or {}).get(k, 0.0) for k in _PNG_KEYS}
        t = trans_phi2delta(cosmo, np.sqrt(kmesh2.numpy()), kpow)
        phik = _safe_div_t(lin_mesh, t)
        phi = irf(phik)
        phi_pos = rd(phi)
        w = w + f["fNL_bp"] * phi_pos
        pd = phi_pos * delta_pos
        sigma_pd = pd.mean()
        w = w + f["fNL_bpd"] * (pd - sigma_pd)
        w = w + f["fNL_bpd2"] * (phi_pos * delta2_pos - 2 * sigma_pd * delta_pos)
        w = w + f["fNL_bps2"] * phi_pos * shear2_pos
        w = w + f["fNL_bn2p"] * rd(irf(-kmesh2 * phik))
    grads = torch.stack([rd(irf(1j * k * lin_mesh)) for k in kvec], dim=-1)
    gcol = g.reshape(-1, 1) if isinstance(g, torch.Tensor) and g.dim() > 0 else g
    dvel = b["bnpar"] * grads * gcol
    return (w, dvel) if png_type is None else (w, dvel, phi)


# ----------------------------------------------------------------------------------------------------------------
# the general evolve (model.py:683-837, 'lpt' / 'nbody' with Lagrangian bias): frames, lines of sight, light cone,
# redshift-space distortions and Alcock-Paczynski of bricks.py:628-877 restated in float64
# ----------------------------------------------------------------------------------------------------------------
def _rot(x, rot, inverse=False):
    if rot is None:
        return x
    m = O._t(rot.as_matrix() if hasattr(rot, "as_matrix") else np.asarray(rot))
    return x @ (m if inverse else m.T)


def cell2phys_pos(pos, center, rot, box, shape):
    """bricks.py:628-636."""
    return _rot(pos * O._t(np.divide(box, shape)) - O._t(np.asarray(box)) / 2, rot) + O._t(np.asarray(center))


def phys2cell_pos(pos, center, rot, box, shape):
    """bricks.py:638-646."""
    return (_rot(pos - O._t(np.asarray(center)), rot, True) + O._t(np.asarray(box)) / 2) / O._t(np.divide(box, shape))


def los_scalefactor_pos(pos, center, rot, box, shape, cosmo, a_obs=None, curved_sky=True):
    """bricks.py:747-766."""
    pos = cell2phys_pos(pos, center, rot, box, shape)
    if curved_sky:
        rpos = pos.norm(dim=-1, keepdim=True)
        los = _safe_div_t(pos, rpos)
    else:
        c = np.asarray(center, dtype=np.float64)
        n = np.linalg.norm(c)
        los = O._t(c / n if n != 0 else np.zeros(3))
        rpos = (pos * los).sum(-1, keepdim=True).abs()
    return los, (O.chi2a(cosmo, rpos) if a_obs is None else a_obs)


def rsd(cosmo, vel, los, a, rot, box, shape, dvel=0.0):
    """bricks.py:781-792."""
    vel = _rot(vel * O._t(np.divide(box, shape)), rot) * (O.a2g(cosmo, a) * O.a2f(cosmo, a)) + dvel
    return (vel * los).sum(-1, keepdim=True) * los


def ap_param(pos, los, alpha_iso, alpha_ap, curved_sky=True):
    """bricks.py:847-856 with isoap2parperp (730-736) and scale_pos (712-720)."""
    if curved_sky:
        return pos * alpha_iso
    a_par, a_perp = alpha_iso * alpha_ap ** (2 / 3), alpha_iso * alpha_ap ** (-1 / 3)
    par = (pos * los).sum(-1, keepdim=True) * los
    return par * a_par + (pos - par) * a_perp


def ap_auto(pos, los, cosmo, cosmo_fid, curved_sky=True):
    """bricks.py:795-813."""
    rpos = pos.norm(dim=-1, keepdim=True) if curved_sky else (pos * los).sum(-1, keepdim=True).abs()
    return pos * _safe_div_t(O.a2chi(cosmo_fid, O.chi2a(cosmo, rpos)), rpos)


def evolve_general(white, transfer, cosmo, init_shape, box_size, evol_shape=None, ptcl_shape=None, paint_shape=None,
                   box_center=(0.0, 0.0, 0.0), box_rot=None, curved_sky=False, a_obs=1.0, evolution="lpt", n_steps=5,
                   a_start=0.0, lpt_order=2, paint_order=2, interlace_order=2, paint_deconv=True, bias=None, rsd_on=True,
                   ap_fid=None, kernel_type="rectangular"):
    """model.py:683-837 (the 'lpt' / 'nbody' branch with Lagrangian bias, no primordial non-Gaussianity)."""
    init_shape = tuple(init_shape)
    evol_shape = init_shape if evol_shape is None else tuple(evol_shape)
    ptcl_shape = evol_shape if ptcl_shape is None else tuple(ptcl_shape)
    paint_shape = init_shape if paint_shape is None else tuple(paint_shape)
    geo = (box_center, box_rot, box_size)
    init_mesh = torch.fft.rfftn(white) * O._t(transfer)
    if evol_shape != init_shape:
        init_mesh = O.chreshape(init_mesh, O.r2chshape(evol_shape))
    pos = O.regular_pos(evol_shape, ptcl_shape)
    _, a = los_scalefactor_pos(pos, *geo, evol_shape, cosmo, a_obs, curved_sky)
    weights, dvel = lagrangian_bias(cosmo, pos, a, box_size, init_mesh, bias or {}, read_order=1)
    if evolution == "lpt":
        dpos, vel = O.lpt(cosmo, init_mesh, pos, a, lpt_order, 1)
        pos = pos + dpos
    else:
        pos, vel = O.nbody_bf(cosmo, init_mesh, pos, a_start, a, n_steps, paint_order, lpt_order, paint_deconv=False)
        pos, vel = pos[-1], vel[-1]
    los, a = los_scalefactor_pos(pos, *geo, evol_shape, cosmo, a_obs, curved_sky)
    pos = cell2phys_pos(pos, *geo, evol_shape)
    if rsd_on:
        pos = pos + rsd(cosmo, vel, los, a, box_rot, box_size, evol_shape, dvel)
    if ap_fid is not None:
        pos = ap_auto(pos, los, cosmo, ap_fid, curved_sky)
    pos = phys2cell_pos(pos, *geo, init_shape)
    gxy = O.nufft(pos, init_shape, None if paint_shape == init_shape else paint_shape, weights, paint_order,
                  interlace_order, kernel_type, paint_deconv)
    gxy = gxy * float(np.divide(init_shape, ptcl_shape).prod())
    if paint_shape != init_shape:
        gxy = O.chreshape(gxy, O.r2chshape(paint_shape))
    return torch.fft.irfftn(gxy, s=paint_shape)
