#!/usr/bin/env python3
"""oracle/slice_ref.py — cut the arithmetic of the reference's hot path out of /root/reference at BUILD time.

TEST INFRASTRUCTURE ONLY.  The reference as a whole cannot be compiled in this container (GCC 13: no
deducing-this, no <print>; Boost/Eigen/nanobind/arts-cat-data absent), but the function bodies that
carry the path's arithmetic are plain `Numeric` + `std::` code.  This script copies those bodies — found
by an anchor on their first line and closed by brace matching, so nothing depends on a line number — into
oracle/_ref/refslice_gen.cpp (git-ignored: reference text never enters this repository's history), wrapped
by oracle/refslice/template.cpp.in + stub.h + api.inc, which are this repo's own glue and compute nothing.
oracle/Makefile compiles the result with g++ 13 into oracle/_ref/librefslice.so, and
tests/test_refslice_pins.py asserts that oracle/oracle.cpp (the restatement every GPU test is checked
against) agrees with this reference object code BITWISE.

Each slice records the line range it was found at (printed into the generated file and into
oracle/_ref/refslice_manifest.json); `expect` is the range SURVEY.md section 8 / VERDICT cite, and a
mismatch is reported as a warning (the reference moved), never silently accepted as a different text: the
anchors must still match exactly once.
"""
from __future__ import annotations

import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (file under REF, regex of the first line, regex of the line that STARTS the last brace block
#          (None: the block opened by the first line), expected (first, last) line numbers)
# mode "lines": first regex .. last regex, both inclusive, no brace matching.
SLICES = {
    # physics_funcs.cc: constants :19-22, invplanck :153-158, planck :192-197, dplanck_dt :254-263
    "physics_consts": ("src/core/physics/physics_funcs.cc", r"inline constexpr Numeric BOLTZMAN_CONST", r"inline constexpr Numeric SPEED_OF_LIGHT", (19, 22), "lines"),
    "invplanck": ("src/core/physics/physics_funcs.cc", r"Numeric invplanck\(const Numeric& i, const Numeric& f\) \{", None, (153, 158), "block"),
    "planck": ("src/core/physics/physics_funcs.cc", r"Numeric planck\(const Numeric& f, const Numeric& t\) \{", None, (192, 197), "block"),
    "dplanck_dt": ("src/core/physics/physics_funcs.cc", r"Numeric dplanck_dt\(const Numeric& f, const Numeric& t\) \{", None, (254, 263), "block"),
    # rtepack value types and their arithmetic
    "stokvec_struct": ("src/core/rtepack/rtepack_stokes_vector.h", r"struct stokvec final : Vector4 \{", None, (13, 69), "block"),
    "stokvec_ops": ("src/core/rtepack/rtepack_stokes_vector.h", r"//! Addition of two stokvec vectors", r"constexpr stokvec avg\(const stokvec &a, const stokvec &b\) \{", (88, 125), "block"),
    "propmat_struct_ops": ("src/core/rtepack/rtepack_propagation_matrix.h", r"struct propmat final : Vector7 \{", r"constexpr propmat avg\(const propmat &a, const propmat &b\) \{", (12, 109), "block"),
    "muelmat_struct_ops": ("src/core/rtepack/rtepack_mueller_matrix.h", r"struct muelmat final : Matrix44 \{", r"constexpr muelmat inv\(const muelmat &A\) \{", (12, 252), "block"),
    "multitype": ("src/core/rtepack/rtepack_multitype.h", r"constexpr muelmat to_muelmat\(const propmat &k\) \{", r"constexpr stokvec operator\*\(const muelmat &a, const stokvec &b\) \{", (11, 68), "block_skip_decls"),
    # the complex 4x4 matrix of the polarised linprop branch and its mixed products (specmat x propmat, muelmat x specmat, ...)
    "specmat_struct_ops": ("src/core/rtepack/rtepack_spectral_matrix.h", r"struct specmat final : ComplexMatrix44 \{", r"constexpr specmat inv\(const specmat &A\) \{", (12, 242), "block"),
    "multitype_specmat": ("src/core/rtepack/rtepack_multitype.h", r"//! Mutliply a specmat with a muelmat matrix", r"constexpr specmat operator-\(const propmat &pm, const specmat &s\) \{", (120, 353), "block"),
    "real_specmat": ("src/core/rtepack/rtepack_multitype.cc", r"muelmat real\(const specmat &A\) \{", None, (116, 133), "block"),
    "specmat_dawson": ("src/core/rtepack/rtepack_spectral_matrix.cc", r"specmat dawson\(const specmat &A\) \{", None, (6, 25), "block"),
    # tran: declaration, then ctor + operator() :20-150, linsrc + linsrc_deriv :207-447, deriv :558-674
    "tran_struct": ("src/core/rtepack/rtepack_transmission.h", r"struct tran \{", None, (69, 109), "block"),
    "tran_ctor_call": ("src/core/rtepack/rtepack_transmission.cc", r"static constexpr Numeric too_small = 1e-4;", r"muelmat tran::operator\(\)\(\) const noexcept \{", (20, 150), "block"),
    "tran_linsrc": ("src/core/rtepack/rtepack_transmission.cc", r"muelmat tran::linsrc\(\) const noexcept \{", r"muelmat tran::linsrc_deriv\(const propmat &dk,", (207, 447), "block"),
    # rte_option linprop: sqrt of a propagation matrix :872-1002, linsrc_linprop + its derivative :449-556
    "propmat_sqrt": ("src/core/rtepack/rtepack_transmission.cc", r"specmat sqrt\(const propmat &pm\) \{", None, (872, 1002), "block"),
    "tran_linprop": ("src/core/rtepack/rtepack_transmission.cc", r"muelmat tran::linsrc_linprop\(const muelmat &t,", r"muelmat tran::linsrc_linprop_deriv\(const muelmat &lambda,", (449, 556), "block"),
    "tran_deriv": ("src/core/rtepack/rtepack_transmission.cc", r"muelmat tran::deriv\(const muelmat &t,", None, (558, 674), "block"),
    # rte_emission's two recursions (anonymous namespace), rtepack_rtestep.cc:265-372
    "rte_constant_linevo": ("src/core/rtepack/rtepack_rtestep.cc", r"void constant\(stokvec_vector_view &Is,", r"void linevo\(stokvec_vector_view &Is,", (265, 371), "block"),
    # lbl: temperature models (namespace model), line::s / ds_dT, single_shape and its builders
    "tmodel_functions": ("src/core/lbl/lbl_temperature_model.h", r"namespace model \{", None, (36, 282), "block"),
    "line_s": ("src/core/lbl/lbl_data.h", r"\[\[nodiscard\]\] Numeric s\(Numeric T, Numeric Q\) const \{", None, (66, 68), "block"),
    "line_ds_dT": ("src/core/lbl/lbl_data.h", r"\[\[nodiscard\]\] Numeric ds_dT\(Numeric T, Numeric Q, Numeric dQ_dT\) const \{", None, (138, 142), "block"),
    "single_shape_struct": ("src/core/lbl/lbl_lineshape_voigt_lte.h", r"struct single_shape \{", None, (20, 109), "block"),
    "line_strength_calc": ("src/core/lbl/lbl_lineshape_voigt_lte.cpp", r"Complex line_strength_calc\(const Numeric inv_gd,", None, (22, 36), "block"),
    "dline_strength_calc_dVMR_dT": ("src/core/lbl/lbl_lineshape_voigt_lte.cpp", r"Complex dline_strength_calc_dVMR\(const Numeric inv_gd,", r"Complex dline_strength_calc_dT\(const Numeric inv_gd,", (86, 143), "block"),
    "line_center_scaled_gd_builder": ("src/core/lbl/lbl_lineshape_voigt_lte.cpp", r"Numeric line_center_calc\(const line& line, const AtmPoint& atm\) \{", r"struct single_shape_builder \{", (145, 204), "block"),
    "single_shape_ctor": ("src/core/lbl/lbl_lineshape_voigt_lte.cpp", r"single_shape::single_shape\(const SpeciesIsotope& spec,", None, (226, 237), "block"),
    "single_shape_F_dF": ("src/core/lbl/lbl_lineshape_voigt_lte.cpp", r"Complex single_shape::F\(const Complex z_\) \{ return Faddeeva::w\(z_\); \}", r"Complex single_shape::dF\(const Complex z_, const Complex F_\) \{", (239, 268), "block"),
    "single_shape_derivs": ("src/core/lbl/lbl_lineshape_voigt_lte.cpp", r"single_shape::zFdF::zFdF\(const Complex z_\)", r"Complex single_shape::dY\(const Complex ds_dY, const Numeric f\) const \{", (270, 339), "block"),
    # the band sum: window of a frequency under a ByLine cutoff, the sums with and without cutoff, scl(f), the clamp and the
    # accumulation into the propagation matrix
    "number_density": ("src/core/physics/physics_funcs.h", r"constexpr Numeric number_density\(Numeric p, Numeric t\) noexcept \{", None, (54, 56), "block"),
    "band_offset_spans": ("src/core/lbl/lbl_lineshape_voigt_lte.h", r"constexpr std::pair<Index, Index> find_offset_and_count_of_frequency_range\(", r"constexpr auto frequency_spans\(const Numeric cutoff,", (123, 155), "block"),
    "band_shape_struct": ("src/core/lbl/lbl_lineshape_voigt_lte.h", r"struct band_shape \{", None, (158, 365), "block"),
    "band_shape_ctor_sum": ("src/core/lbl/lbl_lineshape_voigt_lte.cpp", r"band_shape::band_shape\(std::vector<single_shape>&& ls, const Numeric cut\)", r"Complex band_shape::operator\(\)\(const Numeric f\) const \{", (428, 436), "block"),
    "band_shape_cut": ("src/core/lbl/lbl_lineshape_voigt_lte.cpp", r"Complex band_shape::operator\(\)\(const ConstComplexVectorView& cut,", r"void band_shape::operator\(\)\(ComplexVectorView cut\) const \{", (591, 608), "block"),
    "computedata_ctor_scl": ("src/core/lbl/lbl_lineshape_voigt_lte.cpp", r"ComputeData::ComputeData\(const ConstVectorView& f_grid,", None, (936, 956), "block"),
    "calculate_accumulate": ("src/core/lbl/lbl_lineshape_voigt_lte.cpp", r"const auto F = com_data.scl\[i\] \* com_data.shape\[i\];", r"pm\[i\] \+= zeeman::scale\(com_data.npm, F\);", (1689, 1691), "lines"),
    "zeeman_scale": ("src/core/lbl/lbl_zeeman.h", r"constexpr Propmat scale\(const Propmat &a, const Complex F\) noexcept \{", None, (432, 440), "block"),
}

# the full microwave absorption models (second translation unit, refslice/template_predef.cpp.in)
SLICES_PREDEF = {
    "pwr98_water": ("src/core/predefined/PWR98.cc", r"void water\(PropmatVector& propmat_clearsky,", None, (40, 242), "block"),
    "pwr98_oxygen": ("src/core/predefined/PWR98.cc", r"void oxygen\(PropmatVector& propmat_clearsky,", None, (297, 434), "block"),
    "mpm89_lineshape_h2o": ("src/core/predefined/MPM89.cc", r"constexpr Numeric MPMLineShapeFunction\(const Numeric gamma,", None, (34, 65), "block"),
    "mpm89_water": ("src/core/predefined/MPM89.cc", r"void water\(PropmatVector& propmat_clearsky,", None, (95, 180), "block"),
    "mpm89_lineshape_o2": ("src/core/predefined/MPM89.cc", r"constexpr Numeric MPMLineShapeO2Function\(const Numeric gamma,", None, (203, 236), "block"),
    "mpm89_oxygen": ("src/core/predefined/MPM89.cc", r"void oxygen\(PropmatVector& propmat_clearsky,", None, (270, 411), "block"),
    # Rosenkranz 2021 / 2022: the shared H2O and O2 line-shape functions, the four table-carrying wrappers, the N2 continuum
    "pwr20xx_h2o_shape": ("src/core/predefined/PWR20xx.cc", r"void compute_h2o\(PropmatVector& propmat_clearsky,", None, (21, 166), "block"),
    "pwr20xx_h2o_2021": ("src/core/predefined/PWR20xx.cc", r"void compute_h2o_2021\(PropmatVector& propmat_clearsky,", None, (169, 381), "block"),
    "pwr20xx_h2o_2022": ("src/core/predefined/PWR20xx.cc", r"void compute_h2o_2022\(PropmatVector& propmat_clearsky,", None, (383, 491), "block"),
    "pwr20xx_o2_shape": ("src/core/predefined/PWR20xx.cc", r"void compute_o2\(PropmatVector& propmat_clearsky,", None, (494, 573), "block"),
    "pwr20xx_o2_2021": ("src/core/predefined/PWR20xx.cc", r"void compute_o2_2021\(PropmatVector& propmat_clearsky,", None, (576, 682), "block"),
    "pwr20xx_o2_2022": ("src/core/predefined/PWR20xx.cc", r"void compute_o2_2022\(PropmatVector& propmat_clearsky,", None, (684, 790), "block"),
    "pwr20xx_n2": ("src/core/predefined/PWR20xx.cc", r"void compute_n2\(PropmatVector& propmat_clearsky,", None, (792, 833), "block"),
    "tre05_lineshape_o2": ("src/core/predefined/TRE05.cc", r"constexpr Numeric MPMLineShapeO2Function\(const Numeric gamma,", None, (37, 70), "block"),
    "tre05_oxygen": ("src/core/predefined/TRE05.cc", r"void oxygen\(PropmatVector& propmat_clearsky,", None, (115, 296), "block"),
    "mpm2020_all": ("src/core/predefined/MPM2020.cc", r"constexpr Index num = 38;", r"void compute\(PropmatVector& propmat_clearsky,", (16, 149), "block"),
    "ell07_compute": ("src/core/predefined/ELL07.cc", r"void compute\(PropmatVector& propmat_clearsky,", None, (39, 188), "block"),
    "mtckd400_radfn": ("src/core/predefined/MT_CKD400.cc", r"Numeric RADFN_FUN\(const Numeric XVI, const Numeric XKT\) noexcept \{", None, (37, 78), "block"),
    "mtckd400_xint": ("src/core/predefined/MT_CKD400.cc", r"constexpr Numeric XINT_FUN\(const Numeric P,", None, (85, 93), "block"),
    "mtckd400_check": ("src/core/predefined/MT_CKD400.cc", r"void check\(const WaterData& data\) \{", None, (95, 99), "block"),
    "mtckd400_foreign": ("src/core/predefined/MT_CKD400.cc", r"void compute_foreign_h2o\(PropmatVector& propmat_clearsky,", None, (102, 177), "block"),
    "mtckd400_self": ("src/core/predefined/MT_CKD400.cc", r"void compute_self_h2o\(PropmatVector& propmat_clearsky,", None, (179, 256), "block"),
    "mtckd430_radfn": ("src/core/predefined/MT_CKD430.cc", r"Numeric RADFN_FUN\(const Numeric XVI, const Numeric XKT\) noexcept \{", None, (37, 78), "block"),
    "mtckd430_xint": ("src/core/predefined/MT_CKD430.cc", r"constexpr Numeric XINT_FUN\(const Numeric P,", None, (85, 93), "block"),
    "mtckd430_check": ("src/core/predefined/MT_CKD430.cc", r"void check\(const WaterData& data\) \{", None, (95, 100), "block"),
    "mtckd430_foreign": ("src/core/predefined/MT_CKD430.cc", r"void compute_foreign_h2o\(PropmatVector& propmat_clearsky,", None, (180, 255), "block"),
    "mtckd430_self": ("src/core/predefined/MT_CKD430.cc", r"void compute_self_h2o\(PropmatVector& propmat_clearsky,", None, (257, 334), "block"),
    "mpm93_nitrogen": ("src/core/predefined/MPM93.cc", r"void nitrogen\(PropmatVector& propmat_clearsky,", None, (33, 73), "block"),
}
SLICES.update(SLICES_PREDEF)

# the line-shape model of a line: per-broadener pressure scaling, the broadener mixing loop and its derivatives, the holder of a
# temperature model (third translation unit, refslice/template_mix.cpp.in)
SLICES_MIX = {
    "tmodel_functions_mix": SLICES["tmodel_functions"],
    "tmodel_data_class": ("src/core/lbl/lbl_temperature_model.h", r"class data \{", None, (284, 343), "block"),
    "lsm_structs": ("src/core/lbl/lbl_lineshape_model.h", r"struct species_model \{", r"struct model \{", (17, 224), "block"),
    "lsm_mixing": ("src/core/lbl/lbl_lineshape_model.cpp", r"#define VARIABLE\(name, PVAR, DPVAR\)", r"std::istream& operator>>\(std::istream& is, species_model& x\) \{", (14, 259), "lines_excl"),
}
SLICES.update(SLICES_MIX)


def _strip_for_braces(line: str) -> str:
    """Drop // comments, string and character literals before counting braces."""
    line = re.sub(r'"(?:\\.|[^"\\])*"', '""', line)
    line = re.sub(r"'(?:\\.|[^'\\])'", "''", line)
    return line.split("//", 1)[0]


def _find_unique(lines, pattern, start=0, what=""):
    rx = re.compile(r"\s*" + pattern)
    hits = [i for i in range(start, len(lines)) if rx.match(lines[i])]
    if len(hits) != 1 and start == 0:
        raise SystemExit(f"slice_ref: anchor {pattern!r} matches {len(hits)} lines in {what} (need exactly 1)")
    if not hits:
        raise SystemExit(f"slice_ref: anchor {pattern!r} not found after line {start + 1} in {what}")
    return hits[0]


def _block_end(lines, first, what):
    """Index of the line on which the brace block opened at/after ``first`` closes (depth evaluated at line ends)."""
    depth, opened = 0, False
    for i in range(first, len(lines)):
        s = _strip_for_braces(lines[i])
        depth += s.count("{") - s.count("}")
        opened = opened or "{" in s
        if opened and depth == 0:
            return i
        if depth < 0:
            break
    raise SystemExit(f"slice_ref: unbalanced braces after line {first + 1} of {what}")


def cut(ref, name):
    rel, first_rx, last_rx, expect, mode = SLICES[name]
    path = os.path.join(ref, rel)
    with open(path) as fh:
        lines = fh.read().split("\n")
    first = _find_unique(lines, first_rx, 0, rel)
    if mode in ("lines", "lines_excl"):
        last = _find_unique(lines, last_rx, first, rel) - (mode == "lines_excl")  # lines_excl: up to the line before the anchor
    else:
        last_start = first if last_rx is None else _find_unique(lines, last_rx, first, rel)
        last = _block_end(lines, last_start, rel)
    text = lines[first:last + 1]
    if mode == "block_skip_decls":
        # drop bodiless declarations inside the range (they name container types that are not part of the slice)
        out, i = [], 0
        while i < len(text):
            if re.match(r"\s*stokvec_vector absvec\(", text[i]):
                while not text[i].rstrip().endswith(";"):
                    i += 1
                i += 1
                continue
            out.append(text[i])
            i += 1
        text = out
    found = (first + 1, last + 1)
    if found != tuple(expect):
        print(f"slice_ref: warning: {name}: found {rel}:{found[0]}-{found[1]}, cited range is {expect[0]}-{expect[1]}", file=sys.stderr)
    return rel, found, text


def generate(ref, out_dir, template, out_name, names):
    with open(os.path.join(HERE, "refslice", template)) as fh:
        tpl = fh.read().split("\n")
    manifest, out, used = {}, [], set()
    for ln in tpl:
        m = re.fullmatch(r"@SLICE (\w+)@", ln.strip())
        if not m:
            out.append(ln)
            continue
        name = m.group(1)
        rel, found, text = cut(ref, name)
        used.add(name)
        manifest[name] = {"file": rel, "first": found[0], "last": found[1], "lines": len(text)}
        out.append(f"//>>> SLICE {name}: {rel}:{found[0]}-{found[1]}")
        out.extend(text)
        out.append(f"//<<< SLICE {name}")
    unused = set(names) - used
    if unused:
        raise SystemExit(f"slice_ref: slices never placed by {template}: {sorted(unused)}")
    with open(os.path.join(out_dir, out_name), "w") as fh:
        fh.write("\n".join(out))
    return manifest


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out_dir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    manifest = generate(ref, out_dir, "template.cpp.in", "refslice_gen.cpp", set(SLICES) - set(SLICES_PREDEF) - set(SLICES_MIX))
    manifest.update(generate(ref, out_dir, "template_predef.cpp.in", "refslice_predef_gen.cpp", set(SLICES_PREDEF)))
    manifest.update(generate(ref, out_dir, "template_mix.cpp.in", "refslice_mix_gen.cpp", set(SLICES_MIX)))
    with open(os.path.join(out_dir, "refslice_manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print(f"slice_ref: {len(manifest)} slices, {sum(v['lines'] for v in manifest.values())} reference lines -> {out_dir}/refslice_gen.cpp, refslice_predef_gen.cpp, refslice_mix_gen.cpp")


if __name__ == "__main__":
    main()
