"""JAMS internal units and physical constants (reference src/jams/helpers/consts.h:29-36).

time ps, field T, energy meV, magnetic moment meV/T.
"""
kHBarIU = 0.6582119569              # meV ps
kBohrMagnetonIU = 0.0578838181      # meV / T
kElectronGFactor = 2.0023193043625
kGyromagneticRatioIU = kElectronGFactor * kBohrMagnetonIU / kHBarIU   # rad / (ps T)
kBoltzmannIU = 0.0861733326         # meV / K
kJoule2meV = 6.24150907e21
kmRyd2meV = 13.605693123

# reference src/jams/core/units.h:15-26
ENERGY_UNITS = {
    "joules": kJoule2meV, "J": kJoule2meV,
    "milli_electron_volts": 1.0, "meV": 1.0,
    "milli_rydbergs": kmRyd2meV, "mRyd": kmRyd2meV,
    "rydbergs": kmRyd2meV * 1e3, "Ryd": kmRyd2meV * 1e3,
    "Kelvin": kBoltzmannIU, "K": kBoltzmannIU,
}
