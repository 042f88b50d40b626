"""Voice structures as the host sees them: unit kinds, their I/O ranges and
registers, and the compiler's default autowiring.

Host-side mirror of the reference's struct compilation for the hot-path units
(src/compiler.c:3036-3138 autowiring, src/core.c:163-243 instantiation), so
that callers of the C ABI (include/a2cu.h: a2cu_bank_new takes the wired
chain) describe a voice the way an A2S `struct { ... }` statement does.
"""

GENERATORS = {"wtosc", "fm1", "fm2", "fm3", "fm4", "fm3p", "fm4p", "fm2r", "fm4r"}
FM_OPS = {"fm1": 1, "fm2": 2, "fm3": 3, "fm4": 4, "fm3p": 3, "fm4p": 4,
          "fm2r": 2, "fm4r": 4}
# (mininputs, maxinputs, minoutputs, maxoutputs, matchio)
UNIT_IO = {"panmix": (1, 2, 1, 2, False), "filter12": (1, 2, 1, 2, True),
           "waveshaper": (1, 2, 1, 2, True)}
for _g in GENERATORS:
    UNIT_IO[_g] = (0, 0, 1, 1, False)

REGS = {
    "wtosc": ["w", "p", "a", "phase"],
    "panmix": ["vol", "pan"],
    "filter12": ["cutoff", "q", "lp", "bp", "hp"],
    "waveshaper": ["amount"],
}
for _k, _n in FM_OPS.items():
    r = ["phase", "p", "a", "fb"]
    for _o in range(1, _n):
        r += ["p%d" % _o, "a%d" % _o, "fb%d" % _o]
    REGS[_k] = r

KIND_CODE = {"wtosc": 1, "panmix": 2, "filter12": 3, "waveshaper": 4,
             "fm1": 16, "fm2": 17, "fm3": 18, "fm4": 19, "fm3p": 20,
             "fm4p": 21, "fm2r": 22, "fm4r": 23}


def autowire(kinds, voice_channels=2):
    """Default-I/O autowiring of a struct: compiler.c:3036-3138 followed by the
    instantiation rules of core.c:163-243. Returns a list of
    (kind, ninputs, noutputs, add, wireout)."""
    out = []
    chain = 0
    n = len(kinds)
    for i, k in enumerate(kinds):
        mini, maxi, mino, maxo, matchio = UNIT_IO[k]
        add = 0
        if maxi == 0:
            nin = 0
            if chain:
                add = 1
        else:
            nin = mini
            if not chain:
                raise ValueError("A2_NOINPUT: %s has inputs but no chain" % k)
            if nin != chain:
                # default mininputs == 1; a 2-channel chain needs explicit I/O
                nin = chain
        dsi = any(UNIT_IO[kk][1] > 0 for kk in kinds[i + 1:])
        if i == n - 1 or not dsi:
            wireout = 1
            add = 1
            lo, hi = (nin, nin) if matchio else (mino, maxo)
            nout = min(max(voice_channels, lo), hi)
            chain = 0
        else:
            wireout = 0
            nout = chain if chain else mino
            if matchio:
                nout = nin
            if chain and not nin:
                add = 1
            chain = nout
        out.append((KIND_CODE[k], nin, nout, add, wireout))
    return out
