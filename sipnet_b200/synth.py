"""Deterministic synthetic workloads.

Everything is seeded with SplitMix64 (seed 20260117) so the same arrays come out
on the CPU box and on the GPU box, independent of numpy's Generator version.
There is no network for datasets; the bench and the parity tests at BASELINE
sizes run on these.

The forcing is generated in `.clim` file units (docs/user-guide/model-inputs.md
of the reference) and converted exactly as readClimData() does
(reference src/sipnet/sipnet.c:205-238) by `clim_from_raw`.
"""
from __future__ import annotations

import numpy as np

from . import _abi as A
from .api import SiteData

SEED = 20260117
TINY = 0.000001  # reference src/common/util.h:14

# Base parameter vector: the values of the reference's cropland smoke case
# (tests/smoke/russell_2/sipnet.param), which carries every N-cycle / methane
# parameter.  The two methane rates are scaled by 0.01 for the synthetic base
# (SURVEY 8d caveat: with 0.01 d^-1 soil C collapses within ~200 days and
# relative comparisons become ill-conditioned).
BASE_PARAMS = dict(
    plantWoodInit=2189.40929649864, laiInit=0.0, litterInit=280.0,
    soilInit=2688.13907865276, soilWFracInit=0.261792101162425, snowInit=1.0,
    fineRootFrac=0.2, coarseRootFrac=0.2, aMax=53.2895432752984,
    aMaxFrac=0.898814754704746, psnTMin=-1.09917015364082,
    psnTOpt=12.575202669587, dVpdSlope=0.0311370443404088,
    dVpdExp=1.43241809137722, halfSatPar=26.4931410129647,
    attenuation=0.599056349736876, baseVegResp=0.0329220943439083,
    baseFolRespFrac=0.0322740185147172, baseSoilResp=0.06,
    baseFineRootResp=0.006, baseCoarseRootResp=0.006,
    vegRespQ10=1.48545154678972, fineRootQ10=4.80341959232274,
    coarseRootQ10=4.82091010894416, soilRespQ10=2.9,
    growthRespFrac=0.218386268943034, frozenSoilFolREff=0.0,
    frozenSoilThreshold=0.0, soilRespMoistEffect=1.0, leafOnDay=144.0,
    gddLeafOn=495.32479952033, soilTempLeafOn=12.0, leafOffDay=285.0,
    leafGrowth=114.609399041337, fracLeafFall=0.996631622398798,
    woodTurnoverRate=0.014, leafTurnoverRate=1.02509813978352,
    fineRootTurnoverRate=0.131932290216504, coarseRootTurnoverRate=0.056,
    litterBreakdownRate=0.4, fracLitterRespired=0.5, fineRootAllocation=0.4,
    woodAllocation=0.2, leafAllocation=0.2, waterRemoveFrac=0.088,
    frozenSoilEff=1.0, wueConst=10.9, soilWHC=12.0, immedEvapFrac=0.1,
    leafPoolDepth=0.1, fastFlowFrac=0.0, snowMelt=0.15, rdConst=300.0,
    rSoilConst1=8.2, rSoilConst2=4.3, leafCSpWt=62.4478060526809,
    cFracLeaf=0.45136783760037, minNInit=1.0, soilOrgNInit=135.0,
    litterOrgNInit=14.0, nVolatilizationFrac=0.05, nLeachingFrac=0.25,
    leafCN=20.0, woodCN=100.0, fineRootCN=40.0, kCN=80.0,
    nFixationFracMax=0.5, halfNFixationMax=1.0, fAnoxia=0.7,
    anaerobicDecompRate=0.5, anaerobicTransExp=2.0, soilMethaneRate=0.0001,
    litterMethaneRate=0.0001, waterDrainFrac=1.0, plantStorageNInit=5.0,
    leafNResorptionFrac=0.5, leafOnReallocFrac=0.2, soilCSaturation=1.0,
)
RUSSELL2_METHANE_RATE = 0.01  # member 0 of every ensemble keeps the unmodified russell_2 rates

# Ensemble priors: the "estimated" parameters and [min, max] columns of the
# reference's legacy-format tests/smoke/niwot/sipnet.param.
ENSEMBLE_RANGES = dict(
    aMax=(0.0, 34.0), psnTMin=(-8.0, 8.0), psnTOpt=(5.0, 30.0),
    dVpdSlope=(0.01, 0.25), halfSatPar=(4.0, 27.0),
    baseVegResp=(0.0006, 0.06), baseFolRespFrac=(0.05, 0.3),
    baseFineRootResp=(0.003, 0.6), baseCoarseRootResp=(0.003, 0.6),
    vegRespQ10=(1.4, 2.6), fineRootQ10=(1.4, 5.0), coarseRootQ10=(1.4, 5.0),
    frozenSoilThreshold=(-5.0, 5.0), woodTurnoverRate=(0.001, 1.0),
    leafTurnoverRate=(0.001, 1.0), fineRootTurnoverRate=(0.001, 1.0),
    coarseRootTurnoverRate=(0.001, 1.0), wueConst=(0.01, 109.0),
    soilWHC=(0.1, 36.0), soilWFracInit=(0.0, 1.0), soilRespQ10=(1.4, 5.0),
    baseSoilResp=(0.003, 0.6),
)

# flags of the synthetic configs (C2-C5): litter + anaerobic + nitrogen on top of defaults
SYNTH_FLAGS = dict(A.DEFAULT_FLAGS, litterPool=1, anaerobic=1, nitrogenCycle=1)


class SplitMix64:
    """Vectorised SplitMix64 stream (Steele, Lea, Flood 2014)."""

    def __init__(self, seed: int):
        self.state = np.uint64(seed & 0xFFFFFFFFFFFFFFFF)

    def next_u64(self, n: int) -> np.ndarray:
        with np.errstate(over="ignore"):
            idx = np.arange(1, n + 1, dtype=np.uint64)
            z = self.state + idx * np.uint64(0x9E3779B97F4A7C15)
            self.state = z[-1] if n else self.state
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def uniform(self, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
        u = (self.next_u64(n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        return lo + (hi - lo) * u

    def normal(self, n: int, mu: float = 0.0, sigma: float = 1.0) -> np.ndarray:
        u1 = self.uniform(n)
        u2 = self.uniform(n)
        u1 = np.maximum(u1, 1e-300)
        return mu + sigma * np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)

    def fork(self, k: int) -> "SplitMix64":
        return SplitMix64(int(self.next_u64(1)[0]) ^ (k * 0x632BE59BD9B4E019 & 0xFFFFFFFFFFFFFFFF))


def _is_leap(y: int) -> bool:
    return (y % 4 == 0 and y % 100 != 0) or (y % 400 == 0)


def time_axis(year0: int = 2011, nyears: int = 10, variant: str = "half-daily"):
    """(year, day, time, length) arrays: two records per day.

    half-daily : night at 0.00 and day at 12.00, length 0.5 (T = 7306 for 2011-2020)
    unequal    : niwot-like night 0.417 / day 0.583 d (exercises the mean
                 tracker's partial-eviction branch, runmean.c:76-81)
    """
    years, days, times, lens = [], [], [], []
    for y in range(year0, year0 + nyears):
        nd = 366 if _is_leap(y) else 365
        for d in range(1, nd + 1):
            if variant == "half-daily":
                recs = ((0.0, 0.5), (12.0, 0.5))
            elif variant == "unequal":
                recs = ((0.0, 0.417), (10.0, 0.583))
            else:
                raise ValueError(variant)
            for (tm, ln) in recs:
                years.append(y)
                days.append(d)
                times.append(tm)
                lens.append(ln)
    return (np.array(years, np.int32), np.array(days, np.int32),
            np.array(times, np.float64), np.array(lens, np.float64))


def clim_from_raw(year, day, time, length, tair, tsoil, par, precip, vpd, vpdSoil, vPress, wspd,
                  gdd_flag: int = 1) -> SiteData:
    """The unit conversions / floors of readClimData(), sipnet.c:205-238."""
    length = np.asarray(length, np.float64).copy()
    neg = length < 0
    length[neg] = length[neg] / -86400.0
    tair = np.asarray(tair, np.float64)
    par_c = np.asarray(par, np.float64) * (1.0 / length)
    precip_c = np.asarray(precip, np.float64) * 0.1
    vpd_c = np.asarray(vpd, np.float64) * 0.001
    vpd_c = np.where(vpd_c < TINY, TINY, vpd_c)
    vpdSoil_c = np.asarray(vpdSoil, np.float64) * 0.001
    vPress_c = np.asarray(vPress, np.float64) * 0.001
    wspd_c = np.asarray(wspd, np.float64).copy()
    wspd_c = np.where(wspd_c < TINY, TINY, wspd_c)
    if gdd_flag:
        g = tair * length
        g = np.where(g < 0, 0.0, g)
    else:
        g = np.zeros_like(tair)
    clim = dict(time=np.asarray(time, np.float64), length=length, tair=tair,
                tsoil=np.asarray(tsoil, np.float64), par=par_c, precip=precip_c, vpd=vpd_c,
                vpdSoil=vpdSoil_c, vPress=vPress_c, wspd=wspd_c, gdd=g)
    return SiteData(np.asarray(year, np.int32), np.asarray(day, np.int32), clim)


def synth_site(site_index: int, nyears: int = 10, variant: str = "half-daily", with_events: bool = False,
               gdd_flag: int = 1, year0: int = 2011, seed: int = SEED) -> SiteData:
    """One site's forcing (+ optional C3 event schedule), SURVEY 8d."""
    rng = SplitMix64(seed).fork(1000 + site_index)
    year, day, time, length = time_axis(year0, nyears, variant)
    T = year.size
    is_day = (np.arange(T) % 2) == 1
    doy = day.astype(np.float64)
    Tm = float(rng.uniform(1, 2.0, 18.0)[0])
    season = np.sin(2.0 * np.pi * (doy - 110.0) / 365.0)
    tair = Tm + 12.0 * season + np.where(is_day, 5.0, -5.0) + rng.normal(T, 0.0, 2.0)
    # soil temperature: 10-day lagged, damped air temperature
    lag = 20
    kern = np.ones(lag) / lag
    tair_pad = np.concatenate([np.full(lag - 1, tair[0]), tair])
    tsoil = Tm + 0.6 * (np.convolve(tair_pad, kern, mode="valid") - Tm)
    par = np.where(is_day, rng.uniform(T, 10.0, 45.0) * np.maximum(0.2, 0.5 + 0.5 * season), 0.0)
    wet = rng.uniform(T) < 0.2
    precip = np.where(wet, -6.0 * np.log(np.maximum(rng.uniform(T), 1e-12)), 0.0)
    tfac = np.clip((tair + 10.0) / 40.0, 0.05, 1.0)
    vpd = rng.uniform(T, 50.0, 2500.0) * tfac
    vpdSoil = rng.uniform(T, 50.0, 2500.0) * tfac
    vPress = rng.uniform(T, 300.0, 1800.0)
    wspd = rng.uniform(T, 0.3, 6.0)
    site = clim_from_raw(year, day, time, length, tair, tsoil, par, precip, vpd, vpdSoil, vPress, wspd, gdd_flag)
    if with_events:
        site.events = synth_events(site_index, year0, nyears, seed)
    return site


def synth_events(site_index: int, year0: int = 2011, nyears: int = 10, seed: int = SEED) -> list:
    """Per-year agronomic schedule of config C3: tillage, planting,
    two fertilisations, irrigation every 4 days d150-d240, harvest; days
    jittered +-10 per site."""
    rng = SplitMix64(seed).fork(500000 + site_index)
    jitter = int(np.floor(rng.uniform(1, -10.0, 11.0)[0]))
    ev = []
    for y in range(year0, year0 + nyears):
        ev.append((y, 100 + jitter, A.EV_TILLAGE, 0, 0.2, 0.0, 0.0, 0.0))
        ev.append((y, 120 + jitter, A.EV_PLANTING, 0, 10.0, 3.0, 2.0, 5.0))
        ev.append((y, 121 + jitter, A.EV_FERTILIZATION, 0, 15.0, 5.0, 10.0, 0.0))
        irr = list(range(150 + jitter, 241 + jitter, 4))
        fert2 = 160 + jitter
        merged = sorted([(d, 0) for d in irr] + [(fert2, 1)])
        for d, kind in merged:
            if kind == 0:
                ev.append((y, d, A.EV_IRRIGATION, 1, 2.8, 0.0, 0.0, 0.0))
            else:
                ev.append((y, d, A.EV_FERTILIZATION, 0, 15.0, 5.0, 10.0, 0.0))
        # one canopy irrigation per year reaches the immedEvapFrac split (events.c:488-493)
        ev.append((y, 244 + jitter, A.EV_IRRIGATION, 0, 3.0, 0.0, 0.0, 0.0))
        ev.append((y, 270 + jitter, A.EV_HARVEST, 0, 0.8, 0.0, 0.2, 1.0))
    return ev


def base_param_vector(russell2_methane: bool = False) -> np.ndarray:
    p = np.zeros(A.NPARAMS)
    for k, v in BASE_PARAMS.items():
        p[A.P[k]] = v
    if russell2_methane:
        p[A.P["soilMethaneRate"]] = RUSSELL2_METHANE_RATE
        p[A.P["litterMethaneRate"]] = RUSSELL2_METHANE_RATE
    return p


def synth_params(nmembers: int, stream: int = 0, seed: int = SEED, anchor_member0: bool = True) -> np.ndarray:
    """[80][M] parameter ensemble: base vector with the ENSEMBLE_RANGES entries
    drawn uniformly; member 0 = unmodified russell_2 vector (regression anchor)."""
    rng = SplitMix64(seed).fork(7000000 + stream)
    P = np.repeat(base_param_vector()[:, None], nmembers, axis=1)
    for name in sorted(ENSEMBLE_RANGES):
        lo, hi = ENSEMBLE_RANGES[name]
        P[A.P[name], :] = rng.uniform(nmembers, lo, hi)
    # psnTOpt must sit above psnTMin for a meaningful optimum; keep the draw but
    # repair inverted pairs deterministically (the reference accepts either)
    tmin, topt = P[A.P["psnTMin"]], P[A.P["psnTOpt"]]
    bad = topt <= tmin + 1.0
    topt[bad] = tmin[bad] + 1.0 + (topt[bad] - 5.0) * 0.5
    if anchor_member0 and nmembers > 0:
        P[:, 0] = base_param_vector(russell2_methane=True)
    return np.ascontiguousarray(P)


def synth_obs(nee_member0: np.ndarray, seed: int = SEED) -> np.ndarray:
    """C5 observations: NEE of member 0 + N(0, 0.5^2) noise, 20 % masked (NaN)."""
    rng = SplitMix64(seed).fork(9000001)
    T = nee_member0.size
    obs = nee_member0 + rng.normal(T, 0.0, 0.5)
    obs[rng.uniform(T) < 0.2] = np.nan
    return obs


def config_c2(nmembers: int = 4096, nyears: int = 10, variant: str = "half-daily"):
    """C2: 1 site x nmembers parameter ensemble, full per-step output."""
    site = synth_site(0, nyears, variant)
    return [site], synth_params(nmembers), np.zeros(nmembers, np.int32), dict(SYNTH_FLAGS)


def config_c3(nsites: int = 10000, members_per_site: int = 100, nyears: int = 10, site0: int = 0):
    """C3: nsites synthetic sites x members_per_site, with the events.in schedule."""
    sites = [synth_site(site0 + s, nyears, "half-daily", with_events=True) for s in range(nsites)]
    M = nsites * members_per_site
    params = synth_params(M, stream=1 + site0)
    member_site = np.repeat(np.arange(nsites, dtype=np.int32), members_per_site)
    return sites, params, member_site, dict(SYNTH_FLAGS)
