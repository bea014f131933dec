"""Synthetic VLBI / MRI workloads for the DPI hot path (host-side setup code, numpy only).

The reference builds its operator constants with two un-vendored third-party packages
(`ehtim==1.2.2`, `torchkbnufft==0.3.2.post1`, pinned in /root/reference/DPI.yml:227,271).
Neither is available offline, so the published definitions of the pieces the hot path
consumes are restated here:

* `ftmatrix`      - ehtim.observing.obs_helpers.ftmatrix (called at
                    DPItorch/interferometry_helpers.py:63): dense DFT matrix with a triangle
                    pulse, +2*pi*i sign convention, half-pixel-centred grid.
* `nfft_pulsefac` - ehtim.observing.obs_helpers.NFFTInfo.pulsefac (interferometry_helpers.py:55-56).
* `SyntheticObs`  - a duck-typed stand-in for `ehtim.obsdata.Obsdata` exposing exactly the
                    attributes `Obs_params_torch` touches (interferometry_helpers.py:48,84-90,
                    141-159): `.unpack`, `.data`, `.add_cphase/.cphase`, `.add_camp/.camp`,
                    `.add_logcamp/.logcamp`.

Closure-set *selection* (ehtim's 'min-cut0bl' / 'min' counting) is third-party setup; here a
minimal set per scan is produced directly: triangles through the scan's first station,
(N-1)(N-2)/2 of them, and N(N-3)/2 quadrangles. The hot path is defined on whatever tables
it is handed.
"""
from __future__ import annotations

import numpy as np

RADPERUAS = np.pi / 180.0 / 3600.0 / 1e6  # same constant ehtim exports as eh.RADPERUAS


# ----------------------------------------------------------------------------------------------
# ehtim restatements (published algorithm; parity unpinned - no reference test pins them)
# ----------------------------------------------------------------------------------------------
def triangle_pulse_F(omega, pdim):
    """Fourier-domain 1-D triangle pulse of width `pdim` (ehtim `trianglePulse_F`)."""
    omega = np.asarray(omega, dtype=np.float64)
    out = np.ones_like(omega)
    nz = omega != 0
    out[nz] = (4.0 / (pdim ** 2 * omega[nz] ** 2)) * np.sin(pdim * omega[nz] / 2.0) ** 2
    return out


def triangle_pulse2d_F(omegax, omegay, pdim):
    return triangle_pulse_F(omegax, pdim) * triangle_pulse_F(omegay, pdim)


def ftmatrix(psize, xdim, ydim, uv, pulse=None):
    """Dense DFT matrix [n_vis, ydim*xdim] complex128 (row-major pixel index p=row*xdim+col).

    A[k, p] = pulse_F(2 pi u_k, 2 pi v_k) * exp(2 pi i (x_col u_k + y_row v_k)),
    x_j = (xdim/2 - 1/2 - j) * psize (ehtim convention; consumed transposed at
    interferometry_helpers.py:63-66).
    """
    uv = np.asarray(uv, dtype=np.float64)
    xlist = np.arange(0, -xdim, -1) * psize + (psize * xdim) / 2.0 - psize / 2.0
    ylist = np.arange(0, -ydim, -1) * psize + (psize * ydim) / 2.0 - psize / 2.0
    pf = triangle_pulse2d_F(2 * np.pi * uv[:, 0], 2 * np.pi * uv[:, 1], psize)
    ey = np.exp(2j * np.pi * np.outer(uv[:, 1], ylist))  # [n_vis, ydim]
    ex = np.exp(2j * np.pi * np.outer(uv[:, 0], xlist))  # [n_vis, xdim]
    mat = pf[:, None, None] * ey[:, :, None] * ex[:, None, :]
    return mat.reshape(len(uv), ydim * xdim)


def nfft_pulsefac(psize, uv):
    """NFFTInfo.pulsefac: pulse on the unit grid times the half-image phase shift."""
    uvs = np.asarray(uv, dtype=np.float64) * psize
    phases = np.exp(-1j * np.pi * (uvs[:, 0] + uvs[:, 1]))
    pulses = triangle_pulse2d_F(2 * np.pi * uvs[:, 0], 2 * np.pi * uvs[:, 1], 1.0)
    return pulses * phases


class NFFTInfoLike:
    """Duck-typed NFFTInfo carrying only `.pulsefac` (the single attribute the reference reads)."""

    def __init__(self, xdim, ydim, psize, pulse, npad, p_rad, uv):
        self.xdim, self.ydim, self.psize = xdim, ydim, psize
        self.pulsefac = nfft_pulsefac(psize, uv)


class SimImage:
    """The four attributes of an ehtim Image that Obs_params_torch reads."""

    def __init__(self, npix, fov_uas):
        self.xdim = npix
        self.ydim = npix
        self.psize = fov_uas * RADPERUAS / npix
        self.pulse = triangle_pulse2d_F


# ----------------------------------------------------------------------------------------------
# synthetic observation
# ----------------------------------------------------------------------------------------------
_DT_VIS = [("time", "f8"), ("t1", "U8"), ("t2", "U8"), ("u", "f8"), ("v", "f8"),
           ("vis", "c16"), ("sigma", "f8")]
_DT_CP = [("time", "f8"), ("t1", "U8"), ("t2", "U8"), ("t3", "U8"), ("cphase", "f8"),
          ("sigmacp", "f8")]
_DT_CA = [("time", "f8"), ("t1", "U8"), ("t2", "U8"), ("t3", "U8"), ("t4", "U8"),
          ("camp", "f8"), ("sigmaca", "f8")]


class SyntheticObs:
    """EHT-like observation: `n_stations` sites, `n_scans` scans, a random subset visible per scan.

    Baselines rotate on ellipses in the uv plane (earth-rotation synthesis look-alike) with
    |uv| up to `uvmax` wavelengths. Rows are ordered by scan then by (t1, t2) with t1 < t2 as
    stored names; a seeded fraction of rows is stored with reversed station order so the
    signed-index branches of the closure maps are exercised.
    """

    def __init__(self, n_stations=8, n_scans=100, seed=0, uvmax=8.9e9, min_vis=3, flip_frac=0.25):
        rng = np.random.RandomState(seed)
        names = np.array(["S%02d" % i for i in range(n_stations)])
        # baseline vectors: fixed 3-D station positions projected with hour angle
        pos = rng.normal(size=(n_stations, 3))
        pos /= np.abs(pos).max()
        rows = []
        self.scan_stations = []
        for s in range(n_scans):
            t = 24.0 * s / n_scans
            nvis = rng.randint(min_vis, n_stations + 1)
            vis_st = np.sort(rng.choice(n_stations, size=nvis, replace=False))
            self.scan_stations.append(vis_st)
            ha = 2 * np.pi * t / 24.0
            rot = np.array([[np.cos(ha), -np.sin(ha)], [np.sin(ha), np.cos(ha)]])
            for ia in range(nvis):
                for ib in range(ia + 1, nvis):
                    a, b = vis_st[ia], vis_st[ib]
                    d = pos[a] - pos[b]
                    uvv = rot @ d[:2] * uvmax / 2.0
                    uvv[1] = uvv[1] * 0.8 + 0.2 * d[2] * uvmax / 2.0
                    if rng.rand() < flip_frac:
                        rows.append((t, names[b], names[a], -uvv[0], -uvv[1], 0j, 1.0))
                    else:
                        rows.append((t, names[a], names[b], uvv[0], uvv[1], 0j, 1.0))
        self.data = np.array(rows, dtype=_DT_VIS)
        self.names = names
        self.cphase = None
        self.camp = None
        self.logcamp = None
        self._rng = rng

    # --- the ehtim surface used by Obs_params_torch -------------------------------------------
    def unpack(self, fields):
        out = np.zeros(len(self.data), dtype=[(f, self.data.dtype[f]) for f in fields])
        for f in fields:
            out[f] = self.data[f]
        return out

    def _scans(self):
        times = self.data["time"]
        for s, t in enumerate(np.unique(times)):
            idx = np.where(times == t)[0]
            st = sorted(set(self.data["t1"][idx]) | set(self.data["t2"][idx]))
            yield t, st

    def add_cphase(self, **kw):
        rows = []
        for t, st in self._scans():
            n = len(st)
            for i in range(1, n):
                for j in range(i + 1, n):
                    rows.append((t, st[0], st[i], st[j], 0.0, 1.0))
        self.cphase = np.array(rows, dtype=_DT_CP)

    def _quads(self):
        rows = []
        for t, st in self._scans():
            n = len(st)
            if n < 4:
                continue
            cnt = 0
            want = n * (n - 3) // 2
            for i in range(n):
                for off in range(1, n - 2):
                    if cnt >= want:
                        break
                    q = [st[i], st[(i + off) % n], st[(i + off + 1) % n], st[(i + n - 1) % n]]
                    if len(set(q)) < 4:
                        continue
                    rows.append((t, q[0], q[1], q[2], q[3], 1.0, 1.0))
                    cnt += 1
        return rows

    def add_camp(self, **kw):
        self.camp = np.array(self._quads(), dtype=_DT_CA)

    def add_logcamp(self, **kw):
        self.logcamp = np.array(self._quads(), dtype=_DT_CA)


def ground_truth_image(npix, flux=2.0):
    """Ring + offset blob, total flux `flux` (Jy)."""
    ax = (np.arange(npix) - (npix - 1) / 2.0) / npix
    X, Y = np.meshgrid(ax, ax)
    r = np.hypot(X, Y)
    img = np.exp(-0.5 * ((r - 0.22) / 0.05) ** 2) * (1 + 0.5 * np.cos(np.arctan2(Y, X) - 0.7))
    img += 0.6 * np.exp(-0.5 * (((X - 0.1) / 0.06) ** 2 + ((Y + 0.08) / 0.09) ** 2))
    return img * (flux / img.sum())


def gaussian_prior_image(npix, fov_uas, fwhm_uas, flux):
    """Circular Gaussian prior image (stands in for eh.image.add_gauss at DPI_interferometry.py:101)."""
    sigma = fwhm_uas / (2.0 * np.sqrt(2.0 * np.log(2.0))) / (fov_uas / npix)
    ax = np.arange(npix) - (npix - 1) / 2.0
    X, Y = np.meshgrid(ax, ax)
    img = np.exp(-0.5 * (X ** 2 + Y ** 2) / sigma ** 2)
    return img * (flux / img.sum())


def knee_phantom(npix, flux=0.0907):
    """Knee-shaped ellipse phantom with the flux of the resized fastMRI knee slice."""
    ax = (np.arange(npix) - (npix - 1) / 2.0) / (npix / 2.0)
    X, Y = np.meshgrid(ax, ax)
    img = 0.15 * (((X / 0.8) ** 2 + (Y / 0.9) ** 2) < 1)
    img = img + 0.5 * ((((X + 0.25) / 0.28) ** 2 + ((Y - 0.1) / 0.6) ** 2) < 1)
    img = img + 0.45 * ((((X - 0.3) / 0.25) ** 2 + ((Y + 0.05) / 0.65) ** 2) < 1)
    img = img + 0.3 * np.exp(-((X ** 2 + (Y - 0.6) ** 2) / 0.02))
    k = np.array([0.25, 0.5, 0.25])
    for _ in range(2):  # light separable blur
        img = np.apply_along_axis(lambda m: np.convolve(m, k, mode="same"), 0, img)
        img = np.apply_along_axis(lambda m: np.convolve(m, k, mode="same"), 1, img)
    return (img * (flux / img.sum())).astype(np.float64)


def mri_mask(npix=64, ratio=4, seed=0):
    """Variable-density random Cartesian mask with ~npix^2/ratio samples, centre 16x16 filled and
    fftshifted exactly as DPI_MRI.py:108-111 does with dataset/fastmri_sample/mask/mask4.npy."""
    rng = np.random.RandomState(seed)
    ax = (np.arange(npix) - npix / 2.0) / (npix / 2.0)
    X, Y = np.meshgrid(ax, ax)
    pdf = np.exp(-3.0 * np.hypot(X, Y))
    pdf *= (npix * npix / ratio) / pdf.sum()
    mask = rng.rand(npix, npix) < np.clip(pdf, 0, 1)
    c, h = npix // 2, max(1, npix // 8)  # npix=64: rows/cols 24:40, as DPI_MRI.py:109
    mask[c - h:c + h, c - h:c + h] = 1
    mask = np.fft.fftshift(mask)
    return np.stack((mask, mask), axis=-1).astype(np.float64)
