"""Known-answer tests taken from the reference's own unit tests: the events.in handler
(reference tests/sipnet/test_events_types/testEvent{Harvest,Fertilization,Irrigation,Planting,LeafOnOff,Tillage}.c
and their events_*.in fixtures) and, further down, plant mortality, carbon saturation, methane and drainage
(tests/sipnet/test_modeling/test{PlantMortality,CarbonSaturation,Methane,SoilMoisture,NitrogenCycle}.c).

Those tests call processEvents() + updatePoolsForEvents() on hand-set pools (helpers.c: climate 2024-070,
time 0, length 0.125) and compare the pools with closed-form numbers, tolerance 1e-6.  Our boundary is the
whole step, so each scenario runs ONE complete updateState() with parameters that silence every other flux
(no photosynthesis, respiration, turnover, decomposition, water or nitrogen fluxes); the event handler is then
the only thing that moves a pool, and the reference's expected numbers must come out -- through the oracle, the
live reference (oracle/_ref) and the CUDA kernel alike.  Expected values are the expressions in the reference's
test sources, written out below (`want`); tolerance 1e-9.

Where a reference test starts from a pool of exactly 0 that the full step cannot carry (soil water 0 meets the
TINY floor of calcSoilWaterFluxes, root carbon 0 means plant death, litter/soil carbon 0 makes the N cycle's C:N
ratio 0/0 in the reference itself, nitrogen.c:47-53) the scenario starts from a small positive value instead,
shifts the expected number by it, and says so."""
import numpy as np

from sipnet_b200 import _abi as A, synth
from sipnet_b200.api import SiteData

HARV, FERT, IRRIG, PLANT, TILL, LEAFON, LEAFOFF = (A.EV_HARVEST, A.EV_FERTILIZATION, A.EV_IRRIGATION, A.EV_PLANTING,
                                                   A.EV_TILLAGE, A.EV_LEAFON, A.EV_LEAFOFF)

# every rate that would move a pool without an event
QUIET = dict(aMax=0.0, baseFolRespFrac=0.0, baseVegResp=0.0, baseSoilResp=0.0, woodTurnoverRate=0.0,
             leafTurnoverRate=0.0, fineRootTurnoverRate=0.0, coarseRootTurnoverRate=0.0, baseFineRootResp=0.0,
             baseCoarseRootResp=0.0, litterBreakdownRate=0.0, growthRespFrac=0.0, soilMethaneRate=0.0,
             litterMethaneRate=0.0, nVolatilizationFrac=0.0, nLeachingFrac=0.0, nFixationFracMax=0.0,
             halfNFixationMax=0.0, fastFlowFrac=0.0, snowInit=0.0, immedEvapFrac=0.5, leafOnDay=0.0, leafOffDay=0.0,
             leafGrowth=0.0, fracLeafFall=0.0, soilWHC=1000.0, laiInit=0.0, minNInit=0.0, soilOrgNInit=0.0,
             litterOrgNInit=0.0, plantStorageNInit=0.0, litterInit=0.0, soilInit=0.0, soilWFracInit=0.001)
FLAGS0 = dict(events=1, gdd=0, growthResp=0, leafWater=0, litterPool=0, snow=1, soilPhenol=0, waterHResp=1,
              nitrogenCycle=0, anaerobic=0, flooding=0, carbonSaturation=0)
NFLAGS = dict(FLAGS0, litterPool=1, nitrogenCycle=1, anaerobic=1)


def one_step_site(events, nsteps=1, length=0.125, tsoil=9.0):
    """helpers.c:prepTypesTest(): 2024, day 70, time 0, length 0.125 -- plus benign forcing (no rain, no soil VPD)."""
    year = np.full(nsteps, 2024, np.int32)
    t = np.arange(nsteps) * length
    day = (70 + np.floor(t + 1e-9)).astype(np.int32)
    clim = {k: np.zeros(nsteps) for k in A.CLIM_COLS}
    clim["time"] = np.round((t - np.floor(t + 1e-9)) * 24.0, 6)
    clim["length"][:] = length
    clim["tair"][:] = 10.0
    clim["tsoil"][:] = tsoil
    clim["vpd"][:] = 0.7          # kPa-scale, as left by readClimData
    clim["vpdSoil"][:] = 0.0      # no soil evaporation
    clim["vPress"][:] = 0.7       # above 0.6 kPa: no snow sublimation (sipnet.c:915)
    clim["wspd"][:] = 1.5
    site = SiteData(year, day, clim)
    site.events = [(2024, d, typ, method, p[0], p[1], p[2], p[3]) for (d, typ, method, p) in events]
    return site


def params_for(pools, **over):
    """Parameter vector whose setupModel() yields the given initial pools (sipnet.c:1858-1951)."""
    p = dict(synth.BASE_PARAMS)
    p.update(QUIET)
    leaf, wood, fine, coarse = pools.get("leaf", 0.0), pools.get("wood", 3.0), pools.get("fine", 1.0), pools.get("coarse", 0.0)
    tot = wood + fine + coarse
    p.update(plantWoodInit=tot, fineRootFrac=fine / tot, coarseRootFrac=coarse / tot, laiInit=leaf / p["leafCSpWt"],
             soilInit=pools.get("soilC", 0.0), litterInit=pools.get("litterC", 0.0),
             soilWFracInit=pools.get("water", 1.0) / over.get("soilWHC", p["soilWHC"]), minNInit=pools.get("minN", 0.0),
             snowInit=pools.get("snow", 0.0),
             soilOrgNInit=pools.get("soilOrgN", 0.0), litterOrgNInit=pools.get("litterN", 0.0),
             plantStorageNInit=pools.get("storageN", 0.0))
    p.update(over)
    return np.array([p.get(n, 0.0) for n in A.PARAM_NAMES], dtype=np.float64)


def ev(day, typ, *p, method=0):
    return (day, typ, method, tuple(list(p) + [0.0] * (4 - len(p))))


# name -> (flags, initial pools, parameter overrides, events, expected output columns)
CASES = {}

# ---- testEventHarvest.c: pools leaf 2, wood 3, fine 4, coarse 5, soil 10 (litter 15, orgN 2, litterN 3 with N) ----
_h = dict(leaf=2.0, wood=3.0, fine=4.0, coarse=5.0, soilC=10.0)
CASES["harvest_one_no_litter"] = (FLAGS0, _h, {}, [ev(70, HARV, 0.1, 0.2, 0.3, 0.4)], dict(
    soilC=10 + 0.3 * (2 + 3) + 0.4 * (4 + 5), plantLeafC=2 * (1 - 0.1 - 0.3), plantWoodC=3 * (1 - 0.1 - 0.3),
    fineRootC=4 * (1 - 0.2 - 0.4), coarseRootC=5 * (1 - 0.2 - 0.4)))
_hn = dict(_h, litterC=15.0, soilOrgN=2.0, litterN=3.0, minN=10.0)
_cn = dict(woodCN=10.0, leafCN=20.0, fineRootCN=30.0)
CASES["harvest_two_litter_nitrogen"] = (NFLAGS, _hn, _cn,
                                        [ev(70, HARV, 0.1, 0.2, 0.3, 0.4), ev(70, HARV, 0.2, 0.1, 0.2, 0.1)], dict(
    soilC=10 + (0.4 + 0.1) * (4 + 5), litterC=15 + (0.3 + 0.2) * (2 + 3), plantLeafC=2 * (1 - 0.1 - 0.3 - 0.2 - 0.2),
    plantWoodC=3 * (1 - 0.1 - 0.3 - 0.2 - 0.2), fineRootC=4 * (1 - 0.2 - 0.4 - 0.1 - 0.1),
    coarseRootC=5 * (1 - 0.2 - 0.4 - 0.1 - 0.1), soilOrgN=2 + (4 * (0.4 + 0.1)) / 30.0 + (5 * (0.4 + 0.1)) / 10.0,
    litterN=3 + (3 * (0.3 + 0.2)) / 10.0 + (2 * (0.3 + 0.2)) / 20.0))

# ---- testEventFertilization.c: soil 1.5, litter 1; "fert orgN orgC minN" ----
CASES["fert_one_no_litter_no_nitrogen"] = (FLAGS0, dict(soilC=1.5), {}, [ev(70, FERT, 15, 5, 10)], dict(
    soilC=1 + 5 + 0.5, litterN=0.0, minN=0.0))            # organic C goes to the soil pool; the N cycle is off
CASES["fert_two_litter_nitrogen"] = (NFLAGS, dict(soilC=1.5, litterC=1.0, minN=2.0, litterN=3.0), {},
                                     [ev(70, FERT, 15, 5, 10), ev(70, FERT, 5, 2, 3)], dict(
    litterN=3 + 15 + 5, litterC=1 + 5 + 2, minN=2 + 10 + 3, soilC=1.5))

# ---- testEventIrrigation.c: immedEvapFrac 0.5; method 1 = soil, 0 = canopy (reference starts from soil water 0;
# here 1.0 because a full step floors the water balance at TINY) ----
CASES["irrig_one_soil"] = (FLAGS0, dict(water=1.0), {}, [ev(70, IRRIG, 5, method=1)], dict(soilWater=1 + 5, evapotranspiration=0.0))
CASES["irrig_two_soil_and_canopy"] = (FLAGS0, dict(water=1.0 + 5.0), {},
                                      [ev(70, IRRIG, 3, method=1), ev(70, IRRIG, 4, method=0)],
                                      dict(soilWater=1 + 10, evapotranspiration=2.0))

# ---- testEventPlanting.c: leaf 1, wood 2, fine 3, coarse 4; "plant leafC woodC fineRootC coarseRootC" ----
_p = dict(leaf=1.0, wood=2.0, fine=3.0, coarse=4.0)
CASES["plant_one"] = (FLAGS0, _p, {}, [ev(70, PLANT, 10, 5, 4, 3)],
                      dict(plantLeafC=1 + 10, plantWoodC=2 + 5, fineRootC=3 + 4, coarseRootC=4 + 3))
CASES["plant_two"] = (FLAGS0, _p, {}, [ev(70, PLANT, 10, 5, 4, 3), ev(70, PLANT, 9, 6, 8, 4)],
                      dict(plantLeafC=1 + 19, plantWoodC=2 + 11, fineRootC=3 + 12, coarseRootC=4 + 7))

# ---- testEventLeafOnOff.c: leafGrowth 3, fracLeafFall 0.5, leafCN 30, leafOnReallocFrac 0.5, litter pool on
# (fine roots 1.0 here: the reference's 0 root carbon would be plant death in a full step) ----
_lp = dict(leafGrowth=3.0, fracLeafFall=0.5, leafCN=30.0, leafOnReallocFrac=0.5)
LFLAGS = dict(FLAGS0, litterPool=1)
CASES["leafon_one"] = (LFLAGS, dict(wood=10.0, leaf=0.0), _lp, [ev(70, LEAFON)],
                       dict(plantWoodC=10.0 - 3.0, coarseRootC=0.0, plantLeafC=3.0))
CASES["leafon_two"] = (LFLAGS, dict(wood=10.0, leaf=0.0), _lp, [ev(70, LEAFON), ev(70, LEAFON)],
                       dict(plantWoodC=10.0 - 6.0, coarseRootC=0.0, plantLeafC=6.0))
CASES["leafon_proportional_split"] = (LFLAGS, dict(wood=6.0, coarse=4.0, leaf=0.0), _lp, [ev(70, LEAFON)],
                                      dict(plantWoodC=6.0 - 1.8, coarseRootC=4.0 - 1.2, plantLeafC=3.0))
CASES["leafoff_one_no_nitrogen"] = (LFLAGS, dict(wood=5.0, leaf=10.0), _lp, [ev(70, LEAFOFF)],
                                    dict(plantLeafC=10.0 - 5.0, litterC=5.0, litterN=0.0))
_npools = dict(soilC=10.0, soilOrgN=2.0, litterC=1.0)     # non-empty C pools for the N cycle (see the module docstring)
CASES["leafoff_one_with_nitrogen"] = (NFLAGS, dict(_npools, wood=5.0, leaf=10.0), dict(_lp, leafNResorptionFrac=0.0),
                                      [ev(70, LEAFOFF)], dict(plantLeafC=10.0 - 5.0, litterC=1.0 + 5.0, litterN=5.0 / 30.0))
CASES["leafoff_one_no_litter_pool"] = (FLAGS0, dict(wood=5.0, leaf=10.0), _lp, [ev(70, LEAFOFF)],
                                       dict(plantLeafC=10.0 - 5.0, soilC=5.0, litterN=0.0))
CASES["leafon_carbon_limited"] = (LFLAGS, dict(wood=1.0, leaf=0.0), _lp, [ev(70, LEAFON)],
                                  dict(plantWoodC=1.0 - 0.5, coarseRootC=0.0, plantLeafC=0.5))
_dem = 3.0 / 30.0 - 3.0 / 100.0
CASES["leafon_nitrogen_limited"] = (NFLAGS, dict(_npools, wood=10.0, leaf=0.0, storageN=0.05), dict(_lp, woodCN=100.0), [ev(70, LEAFON)],
                                    dict(plantWoodC=10.0 - 3.0 * (0.05 / _dem), coarseRootC=0.0, plantLeafC=3.0 * (0.05 / _dem)))
_off = 10.0 * 0.5
CASES["leafoff_nitrogen_resorption"] = (NFLAGS, dict(_npools, wood=5.0, leaf=10.0), dict(_lp, leafNResorptionFrac=0.3), [ev(70, LEAFOFF)],
                                        dict(plantLeafC=10.0 - _off, litterC=1.0 + _off, litterN=(_off / 30.0) * (1 - 0.3),
                                             plantStorageN=(_off / 30.0) * 0.3))

# ---- testPlantMortality.c (tests/sipnet/test_modeling): pools wood 5, leaf 2, fine 3, coarse 4, soil 10.  The unit
# test zeroes a pool by hand and calls checkForMortality(); in a whole step the plant has to be alive when the step
# starts (sipnet.c:1828), so death is reached through a total harvest in the same step and the routing of what is
# left follows sipnet.c:1733-1748: roots -> soil C, wood + leaf (+ delta) -> litter C (or soil C), N equivalents ----
_m = dict(wood=5.0, leaf=2.0, fine=3.0, coarse=4.0, soilC=10.0)
_dead = dict(plantWoodC=0.0, plantLeafC=0.0, fineRootC=0.0, coarseRootC=0.0, nppStorage=0.0)
CASES["mortality_roots_harvested_no_litter"] = (FLAGS0, _m, {}, [ev(70, HARV, 0.0, 1.0, 0.0, 0.0)],
                                                dict(_dead, soilC=10.0 + 0.0 + 5.0 + 2.0))
CASES["mortality_shoots_harvested_no_litter"] = (FLAGS0, _m, {}, [ev(70, HARV, 1.0, 0.0, 0.0, 0.0)],
                                                 dict(_dead, soilC=10.0 + 3.0 + 4.0))
CASES["mortality_shoots_harvested_with_litter"] = (LFLAGS, dict(_m, litterC=5.0), {}, [ev(70, HARV, 1.0, 0.0, 0.0, 0.0)],
                                                   dict(_dead, soilC=10.0 + 3.0 + 4.0, litterC=5.0))
_mn = dict(_m, litterC=5.0, soilOrgN=2.0, litterN=3.0, storageN=0.5)
_mcn = dict(woodCN=100.0, leafCN=20.0, fineRootCN=40.0)
CASES["mortality_shoots_harvested_with_nitrogen"] = (NFLAGS, _mn, _mcn, [ev(70, HARV, 1.0, 0.0, 0.0, 0.0)],
                                                     dict(_dead, soilC=17.0, litterC=5.0, soilOrgN=2.0 + 3.0 / 40.0 + 4.0 / 100.0,
                                                          litterN=3.0 + 0.5, plantStorageN=0.0))
CASES["mortality_roots_harvested_with_nitrogen"] = (NFLAGS, _mn, _mcn, [ev(70, HARV, 0.0, 1.0, 0.0, 0.0)],
                                                    dict(_dead, soilC=10.0, litterC=5.0 + 5.0 + 2.0, soilOrgN=2.0,
                                                         litterN=3.0 + 5.0 / 100.0 + 2.0 / 20.0 + 0.5, plantStorageN=0.0))

# ---- testCarbonSaturation.c: soilCSaturation 10, litter 10, a coarse-root loss of 100 or 200 g C m-2 d-1 (here:
# 50 g of coarse roots turning over at 2 or 4 d-1) is split between soil and litter by unitClip(soilC / saturation);
# the 125 % case also respires 50 g C m-2 d-1 from the soil (sipnet.c:1634-1680) ----
CFLAGS = dict(FLAGS0, litterPool=1, carbonSaturation=1)


def _csat(soil, loss, rsoil=0.0):
    sat = min(max(soil / 10.0, 0.0), 1.0)
    return dict(soilC=soil + (loss * (1 - sat) - rsoil) * 0.125, litterC=10.0 + loss * sat * 0.125,
                coarseRootC=50.0 - loss * 0.125)


_cs = dict(soilCSaturation=10.0, soilRespQ10=1.0, soilRespMoistEffect=0.0)
CASES["csat_25pct_loss_100"] = (CFLAGS, dict(coarse=50.0, soilC=2.5, litterC=10.0), dict(_cs, coarseRootTurnoverRate=2.0 * 365.0),
                                [], _csat(2.5, 100.0))
CASES["csat_25pct_loss_200"] = (CFLAGS, dict(coarse=50.0, soilC=2.5, litterC=10.0), dict(_cs, coarseRootTurnoverRate=4.0 * 365.0),
                                [], _csat(2.5, 200.0))
CASES["csat_75pct_loss_100"] = (CFLAGS, dict(coarse=50.0, soilC=7.5, litterC=10.0), dict(_cs, coarseRootTurnoverRate=2.0 * 365.0),
                                [], _csat(7.5, 100.0))
CASES["csat_125pct_loss_200_rsoil_50"] = (CFLAGS, dict(coarse=50.0, soilC=12.5, litterC=10.0),
                                          dict(_cs, coarseRootTurnoverRate=4.0 * 365.0, baseSoilResp=4.0 * 365.0), [],
                                          _csat(12.5, 200.0, 50.0))

# ---- testMethane.c: water 7.5 of WHC 10, fAnoxia 0.6, transition exponent 2, Q10 3 at tsoil 20: temperature effect 9,
# moisture effect ((0.75 - 0.6) / 0.4)^2; rates 0.05 (soil) and 0.1 (litter) ----
_mm = ((0.75 - 0.6) / (1 - 0.6)) ** 2
_mp = dict(soilWHC=10.0, soilRespQ10=3.0, fAnoxia=0.6, anaerobicDecompRate=0.5, anaerobicTransExp=2.0,
           soilMethaneRate=0.05, litterMethaneRate=0.1)
MFLAGS = dict(FLAGS0, litterPool=1, anaerobic=1)
CASES["methane_litter_pool"] = (MFLAGS, dict(water=7.5, soilC=15.0, litterC=7.5), _mp, [], dict(
    soilC=15 - 0.75 * 9.0 * _mm * 0.125, litterC=7.5 - 0.75 * 9.0 * _mm * 0.125, ch4=2 * 0.75 * 9.0 * _mm * 0.125), 20.0)
CASES["methane_no_litter_pool"] = (dict(FLAGS0, anaerobic=1), dict(water=7.5, soilC=20.0), _mp, [], dict(
    soilC=20 - 1.0 * 9.0 * _mm * 0.125, ch4=1.0 * 9.0 * _mm * 0.125), 20.0)

# ---- testSoilMoisture.c: drainage of 2 cm above WHC = 10 with flooding on is min(excess * waterDrainFrac,
# excess / length); snow on the ground switches soil evaporation off (sipnet.c:986, 1016-1027) ----
DFLAGS = dict(FLAGS0, flooding=1)
for _frac, _drain in ((2.0, 4.0), (0.5, 1.0), (0.0, 0.0), (20.0, 16.0)):
    CASES["drainage_frac_%g" % _frac] = (DFLAGS, dict(water=12.0, snow=1.0), dict(soilWHC=10.0, waterDrainFrac=_frac, snowMelt=0.0),
                                         [], dict(soilWater=12.0 - _drain * 0.125, snow=1.0))
CASES["drainage_none_at_whc"] = (DFLAGS, dict(water=10.0, snow=1.0), dict(soilWHC=10.0, waterDrainFrac=1.0, snowMelt=0.0), [],
                                 dict(soilWater=10.0, snow=1.0))

# ---- testNitrogenCycle.c: water 5 of WHC 10, soil C 1.5, litter C 1, Q10 2.9 at tsoil 20 (temperature effect 8.41);
# volatilisation = frac * minN * tEffect * (0.05 + 3.8 A (1 - A)) with A = 0 below fAnoxia; leaching = minN *
# min(drainage / WHC, 1) * frac, drainage being the excess over WHC per day (sipnet.c:1022-1027, nitrogen.c:15-40) ----
_np = dict(soilWHC=10.0, soilRespQ10=2.9, soilRespMoistEffect=1.0, leafCN=20.0, woodCN=100.0, fineRootCN=40.0)
_nq = dict(water=5.0, soilC=1.5, litterC=1.0)
_vol = lambda n: 0.1 * n * 2.9 ** 2 * 0.05                                              # noqa: E731
CASES["nitrogen_volatilisation"] = (NFLAGS, dict(_nq, minN=4.0), dict(_np, nVolatilizationFrac=0.1), [],
                                    dict(minN=4.0 - _vol(4.0) * 0.125, n2o=_vol(4.0) * 0.125), 20.0)
CASES["nitrogen_fertilisation_and_volatilisation"] = (NFLAGS, dict(_nq, minN=2.0), dict(_np, nVolatilizationFrac=0.1),
                                                      [ev(70, FERT, 15, 5, 10)],
                                                      dict(minN=2.0 + (10 / 0.125 - _vol(2.0)) * 0.125, litterN=15.0, litterC=1.0 + 5.0), 20.0)
CASES["nitrogen_leaching_partial"] = (NFLAGS, dict(_nq, water=10.0 + 5 * 0.125, minN=1.0), dict(_np, nLeachingFrac=0.5), [],
                                      dict(minN=1.0 - 1.0 * (5 / 10.0) * 0.5 * 0.125, nLeaching=1.0 * 0.5 * 0.5 * 0.125,
                                           soilWater=10.0), 20.0)
CASES["nitrogen_leaching_capped"] = (NFLAGS, dict(_nq, water=10.0 + 20 * 0.125, minN=1.0), dict(_np, nLeachingFrac=0.5), [],
                                     dict(minN=1.0 - 1.0 * 1.0 * 0.5 * 0.125, nLeaching=0.5 * 0.125, soilWater=10.0), 20.0)

# no event at all: the quiet parameters really leave every pool where it was
CASES["quiet_step_moves_nothing"] = (NFLAGS, dict(leaf=2.0, wood=3.0, fine=4.0, coarse=5.0, soilC=10.0, litterC=15.0,
                                                  soilOrgN=2.0, litterN=3.0, minN=10.0, storageN=1.0, water=7.0), {}, [],
                                     dict(plantLeafC=2.0, plantWoodC=3.0, fineRootC=4.0, coarseRootC=5.0, soilC=10.0,
                                          litterC=15.0, soilOrgN=2.0, litterN=3.0, minN=10.0, plantStorageN=1.0, soilWater=7.0))


def build(name):
    flags, pools, over, events, want, *rest = CASES[name]
    return dict(flags), params_for(pools, **over), one_step_site(events, tsoil=rest[0] if rest else 9.0), want


def tillage_case():
    """testEventTillage.c, second part (events_two_tillage.{in,clim}): tillage 0.5 on day 70 and 0.2 on day 75, 14
    half-day steps; d_till_mod is added to when the event fires and decays by exp(-length / 30) every step."""
    site = one_step_site([ev(70, TILL, 0.5), ev(75, TILL, 0.2)], nsteps=14, length=0.5)
    want, mod = [], 0.0
    for t in range(14):
        if t == 0:
            mod += 0.5
        if site.day[t] == 75 and site.clim["time"][t] == 0.0:
            mod += 0.2
        mod *= np.exp(-0.5 / 30.0)
        want.append(mod)
    return dict(NFLAGS), params_for(dict(leaf=2.0, wood=3.0, fine=4.0, coarse=5.0, soilC=10.0, litterC=15.0)), site, want
