"""Seeded synthetic inputs that stand in for the reference's network queries: a TRILEGAL-like
background-population table and a TIC-like stars table.
There is no MAST/Gaia/TRILEGAL access offline, so benchmarks, tests and examples generate these
tables instead of downloading them.  No compute lives here."""
import numpy as np
import pandas as pd


def trilegal_table(path, n=4000, seed=20260101, tmag_lo=9.0):
    """TRILEGAL-v1.6-like CSV with the columns funcs.trilegal_results reads (funcs.py:353-361):
    Mact, logg, logTe, [M/H], TESS, J, H, Ks; two trailer rows are appended because the reader
    drops the last two lines (`[:-2]`)."""
    rng = np.random.default_rng(seed)
    mass = np.clip(rng.lognormal(np.log(0.6), 0.45, n), 0.09, 2.5)
    logg = np.clip(4.9 - 0.45 * np.log10(mass / 0.2) + rng.normal(0, 0.12, n), 3.0, 5.3)
    giants = rng.random(n) < 0.04
    logg[giants] = rng.uniform(2.0, 3.4, giants.sum())
    teff = np.clip(3000 + 2900 * mass ** 0.8 + rng.normal(0, 150, n), 2700, 11000)
    feh = np.clip(rng.normal(-0.15, 0.3, n), -2.5, 0.5)
    tess = tmag_lo + 12.0 * rng.random(n) ** 0.45
    jk = np.clip(0.95 - 0.11 * (teff - 3000) / 1000 + rng.normal(0, 0.03, n), 0.05, 1.1)
    J = tess - 0.55 - 0.6 * jk
    Ks = J - jk
    H = J - 0.75 * jk
    df = pd.DataFrame({"Gc": 1, "logAge": 9.5, "[M/H]": np.round(feh, 2), "m_ini": np.round(mass, 5),
                       "logL": 0.0, "logTe": np.round(np.log10(teff), 4), "logg": np.round(logg, 3),
                       "m-M0": 10.0, "Av": 0.1, "m2/m1": 0.0, "mbol": 5.0,
                       "TESS": np.round(tess, 3), "J": np.round(J, 3), "H": np.round(H, 3),
                       "Ks": np.round(Ks, 3), "Mact": np.round(mass, 3)})
    trailer = pd.DataFrame({c: [np.nan, np.nan] for c in df.columns})
    pd.concat([df, trailer], ignore_index=True).to_csv(path)
    return path


def stars_table(ID, Tmag, Jmag, Hmag, Kmag, mass, rad, Teff, plx, n_neighbours=10, seed=7):
    """Stars DataFrame with the columns calc_probs reads (triceratops.py:715-752): the target in
    row 0 and `n_neighbours` synthetic neighbours of which exactly one has tdepth > 0."""
    rng = np.random.default_rng(seed)
    rows = [dict(ID=ID, Tmag=Tmag, Jmag=Jmag, Hmag=Hmag, Kmag=Kmag, ra=0.0, dec=0.0,
                 mass=mass, rad=rad, Teff=Teff, plx=plx, sep=0.0, PA=0.0,
                 fluxratio=0.985, tdepth=0.0)]
    for j in range(n_neighbours):
        rows.append(dict(ID=ID + 1000 + j, Tmag=Tmag + rng.uniform(2.5, 6.0),
                         Jmag=np.nan, Hmag=np.nan, Kmag=np.nan, ra=0.0, dec=0.0,
                         mass=np.nan, rad=np.nan, Teff=np.nan, plx=np.nan,
                         sep=rng.uniform(15, 120), PA=rng.uniform(0, 360),
                         fluxratio=0.0015 * (n_neighbours + 1 - j) / n_neighbours, tdepth=0.0))
    df = pd.DataFrame(rows)
    # target: the observed depth needs the full aperture flux; one bright-enough neighbour could
    # host it too, the others would need tdepth > 1
    df.loc[0, "tdepth"] = 0.0105
    if n_neighbours > 0:
        df.loc[1, "tdepth"] = 0.62
        df.loc[1, "fluxratio"] = 0.015
        df.loc[1, ["mass", "rad", "Teff"]] = [0.62, 0.60, 4100.0]
    return df
