"""Pure-Python restatement of the reference's HITRAN .par reader (TEST INFRASTRUCTURE, like the rest of the oracle):
read_par_line (src/core/lbl/lbl_hitran.cpp:66-89), read_hitran_par (:146-172) and hitran_record::from with either
HitranLineStrengthOption (:180-237; S: line::hitran_a lbl_data.cpp:155-169 with einstein_a :34-40).  Unit conversions: src/core/util/arts_conversions.h:51,88,136-138,146.
Python floats are IEEE doubles and float() is correctly rounded, like std::from_chars / fast_float."""
C_LIGHT = 299792458.0
H_PLANCK = 6.62607015e-34


def kaycm2freq(x):
    return x * (100 * C_LIGHT)


GAMMA = kaycm2freq(1 / 101325.0)  # kaycm_per_atm2hz_per_pa(x) = x * kaycm2freq(pa2atm(1))
ENERGY = kaycm2freq(H_PLANCK)     # kaycm2joule(x) = x * kaycm2freq(h)


class HitranError(Exception):
    pass


def _num(s, conv):
    t = s.strip(" ")
    try:
        if conv is int and not t.lstrip("-").isdigit():
            raise ValueError
        if conv is float and (t == "" or t[0] == "+" or any(ch in t for ch in "_ \t") or t.lower() in ("infinity",)):
            raise ValueError
        return conv(t)
    except ValueError:
        raise HitranError(f'Failed to parse value from string "{s}"')


K_BOLTZ = 1.380649e-23
S_FACTOR = kaycm2freq(1e-4)  # kaycm_per_cmsquared2hz_per_msquared


def read_par(text, fmin, fmax, table, option="A"):
    """table: list of (M, I, species, mass[, hitran ratio, Q(296 K)]).  Returns a list of dicts in file order (after the
    window)."""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    out = []
    for ln in lines:
        if len(ln) < 15:
            raise HitranError("Unexpected end of string")
        M, I = _num(ln[0:2], int), ln[2]
        f0 = kaycm2freq(_num(ln[3:15], float))
        if f0 < fmin:
            continue
        isot = next((k for k, t in enumerate(table) if t[0] == M and t[1] == I), None)
        if isot is None:
            raise HitranError("not in the isotopologue table")
        if len(ln) < 160:
            raise HitranError("Unexpected end of string")
        S = _num(ln[15:25], float)  # only used with option S
        rec = dict(isot=isot, f0=f0, a=_num(ln[25:35], float), gamma_air=_num(ln[35:40], float) * GAMMA,
                   gamma_self=_num(ln[40:45], float) * GAMMA, e0=_num(ln[45:55], float) * ENERGY, n=_num(ln[55:59], float),
                   delta=_num(ln[59:67], float) * GAMMA, gu=_num(ln[146:153], float), gl=_num(ln[153:160], float))
        if len(ln) > 161:  # one separator character after the par block is skipped (lbl_hitran.cpp:126), then :133-135
            raise HitranError("Part of the line was not parsed")
        import math
        if option == "S":
            if rec["gu"] == 0.0:
                rec["gu"] = rec["gl"] = -1.0
            s = (S * S_FACTOR) / table[isot][4]
            cf = C_LIGHT / f0
            T0 = 296.0
            rec["a"] = -8.0 * math.pi * table[isot][5] * s / (
                rec["gu"] * math.exp(-rec["e0"] / (K_BOLTZ * T0)) * math.expm1(-(H_PLANCK * f0) / (K_BOLTZ * T0)) * (cf * cf))
        if not (math.isfinite(rec["a"]) and abs(rec["a"]) >= 2.2250738585072014e-308) or \
           not (math.isfinite(rec["gu"]) and abs(rec["gu"]) >= 2.2250738585072014e-308):
            raise HitranError("Invalid Einstein coefficient")
        if f0 > fmax:
            break
        out.append(rec)
    return out
