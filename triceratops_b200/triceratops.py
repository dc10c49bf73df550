"""`target.calc_probs(...)`: the caller of the marginal-likelihood path, as a drop-in.

Mirrors the reference's triceratops/triceratops.py:673-1485 (calc_probs) with the same signature,
scenario table layout (15 target-star rows + 3 per nearby star with tdepth > 0), attributes
(`probs`, `lnZ`, `FPP`, `NFPP`, `FPP_degenerate`, `star_num`, `u1`, `u2`, `fluxratio_EB`,
`fluxratio_comp`) and warnings.

`calc_depths` (triceratops.py:559-671), which produces the `fluxratio` / `tdepth` columns
calc_probs reads, is included as host numpy.  The rest of the reference's `target` class --
TIC/Gaia/TessCut queries in __init__, plotting, star-table bookkeeping -- is outside this
package's scope: there is no network here, so a `target` is built
from a stars table the caller already has (the reference's `target.stars` DataFrame, e.g. saved
from an online session), optionally its `pix_coords`, and a saved TRILEGAL file.
"""
import collections
import functools
import sys
import warnings

import numpy as np
from pandas import DataFrame
from scipy.special import ndtr

from . import _bufpool, _dispatch
from ._numerics import _normalize_probabilities
from .funcs import renorm_flux
from .marginal_likelihoods import (lnZ_BEB, lnZ_BTP, lnZ_DEB, lnZ_DTP, lnZ_PEB, lnZ_PTP, lnZ_SEB,
                                   lnZ_STP, lnZ_TEB, lnZ_TTP)

_STAR_COLUMNS = ("ID", "Tmag", "Jmag", "Hmag", "Kmag", "ra", "dec", "mass", "rad", "Teff", "plx",
                 "fluxratio", "tdepth")

# scenario rows whose results may still be in flight while the next scenario is being drawn
_PIPELINE_DEPTH = 3

_RESULT_KEYS = ("M_s", "R_s", "u1", "u2", "P_orb", "inc", "b", "R_p", "ecc", "argp", "M_EB",
                "R_EB", "fluxratio_EB", "fluxratio_comp")

# (drop_scenario key, first row, star_num, scenario names) in the reference's order
# (triceratops.py:784-1332)
_TARGET_SCENARIOS = (
    ("TP", 0, 1, ("TP",)),
    ("EB", 1, 1, ("EB", "EBx2P")),
    ("PTP", 3, 1, ("PTP",)),
    ("PEB", 4, 1, ("PEB", "PEBx2P")),
    ("STP", 6, 2, ("STP",)),
    ("SEB", 7, 2, ("SEB", "SEBx2P")),
    ("DTP", 9, 1, ("DTP",)),
    ("DEB", 10, 1, ("DEB", "DEBx2P")),
    ("BTP", 12, 2, ("BTP",)),
    ("BEB", 13, 2, ("BEB", "BEBx2P")),
)


class target:
    def __init__(self, ID: int, sectors=None, search_radius: int = 10, mission: str = "TESS",
                 lightkurve_cache_dir=None, trilegal_fname=None, ra: float = None,
                 dec: float = None, verify_ssl: bool = True, stars: DataFrame = None,
                 pix_coords=None):
        """Offline constructor: same leading arguments as the reference (triceratops.py:42-45)
        plus `stars`, the table the reference would have assembled from TIC (one row per star,
        target first, columns ID, Tmag, Jmag, Hmag, Kmag, ra, dec, mass, rad, Teff, plx,
        fluxratio, tdepth).  `pix_coords` (optional, one [n_stars, 2] array of pixel positions
        per sector, the reference's `target.pix_coords`) enables `calc_depths`."""
        if mission != "TESS" and mission != "Kepler" and mission != "K2":
            raise ValueError("Introduced invalid mission: " + mission)
        if stars is None:
            raise NotImplementedError(
                "catalogue queries are outside triceratops_b200 (no network): pass the stars "
                "table as target(..., stars=DataFrame)")
        stars = stars.copy()
        if pix_coords is not None:
            for c in ("fluxratio", "tdepth"):       # filled by calc_depths
                if c not in stars.columns:
                    stars[c] = 0.0
        missing = [c for c in _STAR_COLUMNS if c not in stars.columns]
        if missing:
            raise ValueError("stars table lacks columns: " + ", ".join(missing))
        self.pix_coords = None if pix_coords is None else [np.asarray(p, float)
                                                           for p in pix_coords]
        self.ID = ID
        self.mission = mission
        self.sectors = sectors
        self.search_radius = search_radius
        self.N_pix = 2 * search_radius + 2
        self.stars = stars.reset_index(drop=True)
        self.trilegal_fname = trilegal_fname
        self.trilegal_url = None

    def calc_depths(self, tdepth: float, all_ap_pixels=None):
        """Flux ratio of every star inside the photometric aperture(s) and the transit depth
        each would need to cause the observed one (reference triceratops.py:559-671): circular
        Gaussian PSF of 0.75 px integrated analytically over each aperture pixel, averaged over
        apertures.  Fills stars["fluxratio"] and stars["tdepth"], the inputs of calc_probs."""
        if self.pix_coords is None:
            raise RuntimeError("calc_depths needs pix_coords (pass them to target())")
        if all_ap_pixels is None:
            print("No apertures provided, assuming 5x5 centered on target.")
            all_ap_pixels = []
            for pc in self.pix_coords:
                c = np.round(pc[0])
                all_ap_pixels.append(np.array([
                    np.repeat(np.arange(c[0] - 2, c[0] + 3, 1), 5),
                    np.tile(np.arange(c[1] - 2, c[1] + 3, 1), 5)]).T)
        sigma = 0.75
        Tmag = self.stars.Tmag.values
        A = 10 ** ((np.min(Tmag) - Tmag) / 2.5)        # flux relative to the brightest star
        ratios = np.zeros([len(all_ap_pixels), len(self.stars)])
        for k, pixels in enumerate(all_ap_pixels):
            pixels = np.array(pixels)
            mu = self.pix_coords[k]
            # separable pixel-box integral: [stars, pixels]
            gx = (ndtr((pixels[None, :, 0] + 0.5 - mu[:, None, 0]) / sigma)
                  - ndtr((pixels[None, :, 0] - 0.5 - mu[:, None, 0]) / sigma))
            gy = (ndtr((pixels[None, :, 1] + 0.5 - mu[:, None, 1]) / sigma)
                  - ndtr((pixels[None, :, 1] - 0.5 - mu[:, None, 1]) / sigma))
            rel = np.array([A[i] * np.sum(gx[i] * gy[i]) for i in range(len(A))])
            ratios[k, :] = rel / np.sum(rel)
        flux_ratios = np.mean(ratios, axis=0)
        self.stars["fluxratio"] = flux_ratios
        tdepths = np.zeros(len(self.stars))
        nz = flux_ratios != 0
        tdepths[nz] = 1 - (flux_ratios[nz] - tdepth) / flux_ratios[nz]
        tdepths[tdepths > 1] = 0
        self.stars["tdepth"] = tdepths
        hosts = self.stars[self.stars["tdepth"] > 0]
        for i, ID in enumerate(hosts["ID"].values):
            need = ["mass", "rad", "Teff"] + (["plx"] if i == 0 else [])
            if any(np.isnan(hosts[c].values[i]) for c in need):
                print("WARNING: " + str(ID) + " is missing stellar properties"
                      + (" required for validation." if i == 0
                         else "; Solar values will be assumed."))
        return

    def calc_probs(self, time: np.ndarray, flux_0: np.ndarray,
                   flux_err_0: float, P_orb,
                   contrast_curve_file: str = None, filt: str = "TESS",
                   N: int = 1000000, parallel: bool = False,
                   drop_scenario: list = [],
                   verbose: int = 1, flatpriors: bool = False,
                   exptime: float = 0.00139, nsamples: int = 20,
                   molusc_file: str = None):
        """Relative probability of every scenario, FPP and NFPP (triceratops.py:673-1485)."""
        keep = ~np.isnan(time) & ~np.isnan(flux_0)
        time = time[keep]
        flux_0 = flux_0[keep]
        if np.ndim(flux_err_0) != 0:
            # an error per time stamp (an extension of the reference's scalar flux_err_0, see
            # tri_set_lightcurve_err): chi^2 is weighted per point, the scalar formulas use the mean
            flux_err_0 = np.asarray(flux_err_0, dtype=float)[keep]
        filtered = self.stars[self.stars["tdepth"] > 0]
        n_rows = 3 * len(filtered) + 12
        targets = np.zeros(n_rows, dtype=np.dtype("i8"))
        star_num = np.zeros(n_rows, dtype=np.dtype("i8"))
        scenarios = np.zeros(n_rows, dtype=np.dtype("U6"))
        best = {k: np.zeros(n_rows) for k in _RESULT_KEYS}
        lnZ = np.zeros(n_rows)

        # Results are read a few scenarios after they were requested: the GPU works on one
        # scenario while the host draws the priors of the next (_dispatch.deferring), and with
        # numpy's draws (host sampler) the scenario functions run in a few threads chained so
        # that the global generator is still consumed in the reference's order
        # (_dispatch.ScenarioChain).
        waiting = collections.deque()      # (rows, pending scenario call)
        held = []                          # (rows, results) waiting for the group's exchange
        from . import marginal_likelihoods as _ml
        host_sampler = _ml._sampler_mode() == "host"
        threads = _dispatch.scenario_threads(_dispatch._dist() is not None) if host_sampler else 1
        eng = _dispatch.get_engine()   # created here, not by whichever scenario thread comes first
        chain = _dispatch.ScenarioChain(threads)
        # Under a process group (one process per GPU) the evidence records and best-draw
        # candidates of ALL rows are exchanged with one all-gather at the end; with numpy's
        # draws rank 0 alone runs the scenario functions and scatters every engine call's
        # columns, the other ranks evaluate their slices (_dispatch.CallGroup).
        group = _dispatch.open_group(scatter=host_sampler)

        def fill(rows, parts):
            for j, res in zip(rows, parts):
                res = _dispatch.resolve(res)
                for k in _RESULT_KEYS:
                    best[k][j] = res[k][0]
                lnZ[j] = res["lnZ"]

        def settle(keep):
            while len(waiting) > keep:
                rows, call = waiting.popleft()
                out = call.result()
                parts = out if isinstance(out, tuple) else (out,)
                if group is None:
                    fill(rows, parts)
                else:   # the local half now (frees the engine's slot), the rest after the exchange
                    for res in parts:
                        _dispatch.prepare(res)
                    held.append((rows, parts))

        def launch(row, ID, num, names, fn):
            """Start one lnZ_* call (one or two table rows) and read older results."""
            rows = [row + k for k in range(len(names))]
            for j, name in zip(rows, names):
                targets[j], star_num[j], scenarios[j] = ID, num, name
            if fn is None:
                for j in rows:
                    lnZ[j] = -np.inf
                return
            waiting.append((rows, chain.run(fn)))
            settle(_PIPELINE_DEPTH)

        def say(msg):
            if verbose == 1:
                print(msg)

        if self.trilegal_fname is None:
            raise RuntimeError("no saved TRILEGAL table: pass trilegal_fname to target(); "
                               "the online query of the reference is out of scope")
        trilegal_fname = self.trilegal_fname

        follower = group is not None and group.scatter and not group.is_root
        ok = False
        # the thread that holds numpy's generator hands the GIL back and forth with the
        # scenario threads ~150 times per call: a short switch interval keeps those hand-overs
        # from costing 5 ms each
        switch_interval = sys.getswitchinterval()
        sys.setswitchinterval(_dispatch.gil_switch_interval())
        _bufpool.note_call()       # (page-locked column buffers from a process's second call on)
        try:
            with _dispatch.deferring():
                for i, ID in enumerate(filtered["ID"].values if not follower else ()):
                    star = {c: filtered[c].values[i] for c in _STAR_COLUMNS}
                    flux, flux_err = renorm_flux(flux_0, flux_err_0, star["fluxratio"])
                    M_s, R_s, Teff, plx = star["mass"], star["rad"], star["Teff"], star["plx"]
                    mags = (star["Tmag"], star["Jmag"], star["Hmag"], star["Kmag"])
                    Z = 0.0
                    lc = (time, flux, flux_err, P_orb)
                    tail = (N, parallel, self.mission, flatpriors, exptime, nsamples)

                    if i == 0:
                        if np.isnan(M_s) or np.isnan(R_s) or np.isnan(Teff) or np.isnan(plx):
                            print("Insufficient information to validate " + str(ID)
                                  + ". Please ensure a stellar mass (in M_Sun), radius (in R_Sun), "
                                  + "Teff (in K), and plx (in mas) are provided in the .stars dataframe.")
                            break
                        bind = functools.partial      # (the loop variables change under the threads)
                        runners = {
                            "TP": bind(lnZ_TTP, *lc, M_s, R_s, Teff, Z, *tail),
                            "EB": bind(lnZ_TEB, *lc, M_s, R_s, Teff, Z, *tail),
                            "PTP": bind(lnZ_PTP, *lc, M_s, R_s, Teff, Z, plx, contrast_curve_file,
                                        filt, *tail, molusc_file),
                            "PEB": bind(lnZ_PEB, *lc, M_s, R_s, Teff, Z, plx, contrast_curve_file,
                                        filt, *tail, molusc_file),
                            "STP": bind(lnZ_STP, *lc, M_s, R_s, Teff, Z, plx, contrast_curve_file,
                                        filt, *tail, molusc_file),
                            "SEB": bind(lnZ_SEB, *lc, M_s, R_s, Teff, Z, plx, contrast_curve_file,
                                        filt, *tail, molusc_file),
                            "DTP": bind(lnZ_DTP, *lc, M_s, R_s, Teff, Z, *mags, trilegal_fname,
                                        contrast_curve_file, filt, *tail),
                            "DEB": bind(lnZ_DEB, *lc, M_s, R_s, Teff, Z, *mags, trilegal_fname,
                                        contrast_curve_file, filt, *tail),
                            "BTP": bind(lnZ_BTP, *lc, M_s, R_s, Teff, *mags, trilegal_fname,
                                        contrast_curve_file, filt, *tail),
                            "BEB": bind(lnZ_BEB, *lc, M_s, R_s, Teff, *mags, trilegal_fname,
                                        contrast_curve_file, filt, *tail),
                        }
                        for key, row, num, names in _TARGET_SCENARIOS:
                            if len(names) == 1:
                                say("Calculating " + names[0] + " scenario probability for "
                                    + str(ID) + ".")
                            else:
                                say("Calculating " + names[0] + " and " + names[1]
                                    + " scenario probabilities for " + str(ID) + ".")
                            launch(row, ID, num, names,
                                   None if key in drop_scenario else runners[key])
                    else:
                        # nearby star: unknown properties default to solar (triceratops.py:1345-1350)
                        if np.isnan(Teff):
                            Teff = 5777
                        if np.isnan(M_s):
                            M_s = 1.0
                        if np.isnan(R_s):
                            R_s = 1.0
                        say("Calculating NTP, NEB, and NEB2xP scenario probabilities for "
                            + str(ID) + ".")
                        row = 15 + 3 * (i - 1)
                        launch(row, ID, 1, ("NTP",),
                               functools.partial(lnZ_TTP, *lc, M_s, R_s, Teff, Z, *tail))
                        launch(row + 1, ID, 1, ("NEB", "NEBx2P"),
                               functools.partial(lnZ_TEB, *lc, M_s, R_s, Teff, Z, *tail))
                settle(0)
                ok = True
        finally:
            chain.close()
            sys.setswitchinterval(switch_interval)
            try:
                if group is not None and group.scatter and group.is_root:
                    group.finish_root(error=not ok)
            finally:
                if not ok:
                    _dispatch.close_group()
        try:
            if follower:
                targets, star_num, scenarios, best, lnZ = group.follow(eng)
            elif group is not None:
                group.exchange()
                for rows, parts in held:
                    fill(rows, parts)
                if group.scatter:
                    group.publish((targets, star_num, scenarios, best, lnZ))
        finally:
            self.collectives = None if group is None else {
                "record_exchanges": group.collectives, "scatters": group.scatters}
            _dispatch.close_group()

        relative_probs, status = _normalize_probabilities(lnZ)
        if status == 'anomaly':
            warnings.warn(
                "Unexpected NaN or +inf in scenario log-evidences. This indicates a numerical "
                "anomaly unrelated to geometric exclusions. Inspect self.lnZ for diagnostics.",
                RuntimeWarning, stacklevel=2)
            self.FPP_degenerate = True
        elif status == 'all_neginf':
            warnings.warn(
                "All scenario log-evidences are -inf: every MC draw was geometrically invalid. "
                "FPP=1.0 reflects a failed computation, not a confident false positive. "
                "Inspect self.lnZ for diagnostics.",
                RuntimeWarning, stacklevel=2)
            self.FPP_degenerate = True
        else:
            self.FPP_degenerate = False

        self.probs = DataFrame({
            "ID": targets, "scenario": scenarios,
            "M_s": best["M_s"], "R_s": best["R_s"], "P_orb": best["P_orb"], "inc": best["inc"],
            "b": best["b"], "ecc": best["ecc"], "w": best["argp"], "R_p": best["R_p"],
            "M_EB": best["M_EB"], "R_EB": best["R_EB"], "prob": relative_probs,
        })
        self.lnZ = lnZ
        self.star_num = star_num
        self.u1 = best["u1"]
        self.u2 = best["u2"]
        self.fluxratio_EB = best["fluxratio_EB"]
        self.fluxratio_comp = best["fluxratio_comp"]
        prob = self.probs.prob
        self.FPP = 1 - (prob[0] + prob[3] + prob[9])
        self.NFPP = np.sum(prob[15:]) if len(prob) > 15 else 0.0
        return
