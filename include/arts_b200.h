/* arts_b200.h — C ABI of the B200-native clear-sky spectral hot path for ARTS.
 *
 * This is the drop-in boundary: the bodies of the ARTS workspace methods
 *   spectral_propmatAddLines             (reference src/m_lbl.cc:242-300)
 *   spectral_propmat_pathFromPath        (reference src/m_propmat.cc:5-65)
 *   spectral_tramat_pathFromPath         (reference src/m_tramat.cc:3-27)
 *   spectral_rad_srcvec_pathFromPropmat  (reference src/m_srcvec.cc:8-31)
 *   spectral_radStepByStepEmission       (reference src/m_spectral_radiance.cc:18-46)
 * become shims that flatten their C++ arguments into the plain arrays below
 * and call these functions.  No torch / C++ types cross
 * the boundary.  Every function returns 0 on success; on failure it returns a
 * non-zero AB200_ERR_* code and ab200_last_error() holds the message (thread
 * local).  There is NO CPU fallback: without a usable CUDA device every
 * compute entry point fails with AB200_ERR_CUDA.
 *
 * All floating point data are IEEE double ("Numeric", reference
 * src/core/util/configtypes.h:7-13).  Layouts of Propmat / Stokvec / Muelmat
 * arrays are the reference's own storage layouts (row-major contiguous,
 * src/core/matpack/matpack_mdspan_data_t.h:38-47):
 *   Propmat  = 7 doubles [A,B,C,D,U,V,W]   (rtepack_propagation_matrix.h:12-30)
 *   Stokvec  = 4 doubles [I,Q,U,V]
 *   Muelmat  = 16 doubles, row-major 4x4
 */
#ifndef ARTS_B200_H
#define ARTS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ------------------------------------------------------ */
#define AB200_OK 0
#define AB200_ERR_INVALID 1     /* bad argument or shape (the shim's ARTS_USER_ERROR) */
#define AB200_ERR_UNSUPPORTED 2 /* input outside the path (non VP_LTE band, pressure target, ...) */
#define AB200_ERR_CUDA 3        /* CUDA runtime failure / no device */
#define AB200_ERR_NOMEM 4

/* ---- enums mirrored from the reference -------------------------------- */
/* SpeciesEnum::Bath: "all other species" broadener, and "all species" as
 * select_species (lbl_lineshape.cpp:191). */
#define AB200_SPECIES_BATH (-1)

/* LineShapeModelVariable subset used by the Voigt LTE engine
 * (lbl_lineshape_voigt_lte.cpp:22-36,145-204). */
enum { AB200_VAR_G0 = 0, AB200_VAR_D0 = 1, AB200_VAR_DV = 2, AB200_VAR_Y = 3, AB200_VAR_G = 4, AB200_NVAR = 5 };

/* LineShapeModelType (lbl_temperature_model.h:18-34); POLY limited to 4 coefficients. */
enum {
  AB200_TM_ABSENT = -1,
  AB200_TM_T0 = 0,
  AB200_TM_T1 = 1,
  AB200_TM_T2 = 2,
  AB200_TM_T3 = 3,
  AB200_TM_T4 = 4,
  AB200_TM_T5 = 5,
  AB200_TM_AER = 6,
  AB200_TM_DPL = 7,
  AB200_TM_POLY = 8
};

/* LineByLineLineshape: only VP_LTE is on the path. */
/* LineByLineLineshape: VP_LTE (engine A) and VP_LTE_MIRROR (lbl_lineshape_voigt_lte_mirrored.cpp: every line plus its
 * mirror image at -f0) are on the path; anything else is AB200_ERR_UNSUPPORTED. */
enum { AB200_LINESHAPE_VP_LTE = 0, AB200_LINESHAPE_OTHER = 1, AB200_LINESHAPE_VP_LTE_MIRROR = 2 };
/* LineByLineCutoffType (lbl_data.h:178-194). */
enum { AB200_CUTOFF_NONE = 0, AB200_CUTOFF_BYLINE = 1 };
/* TransmittanceOption (arts_options.cc:953-1030); linsrc is the default rte_option. */
enum { AB200_RTE_CONSTANT = 0, AB200_RTE_LINSRC = 1, AB200_RTE_LINPROP = 2 };
/* Jacobian target kinds that reach the kernels (AtmKey::t, SpeciesEnum VMR, AtmKey::wind_u/v/w, AtmKey::mag_u/v/w).  The wind rows are the
 * frequency derivative of the line absorption (single_shape::df, lbl_lineshape_voigt_lte.cpp:275, :1036-1062, :1514-1523)
 * times f * freq_wind_shift_jac (spectral_propmat_jacWindFix, src/m_frequency_grid.cc:106-182, wind_shift :56-82). */
enum {
  AB200_TARGET_T = 0,
  AB200_TARGET_VMR = 1,
  AB200_TARGET_WIND_U = 2,
  AB200_TARGET_WIND_V = 3,
  AB200_TARGET_WIND_W = 4,
  /* AtmKey::mag_u/v/w: Zeeman polarisations only (lbl_lineshape_voigt_lte.cpp:1484-1513): the splitting derivative
   * s dz dF with dz = -inv_gd (mag_c / H) Splitting (:1066-1162, single_shape::dH :305-307) through the polarisation
   * matrix and the matrix's own derivative dnorm_view_d{u,v,w} (lbl_zeeman.cpp:361-411, 457-536; scale lbl_zeeman.h:442-453) */
  AB200_TARGET_MAG_U = 5,
  AB200_TARGET_MAG_V = 6,
  AB200_TARGET_MAG_W = 7,
  /* JacobianTargets::line (lbl::line_key, lbl_lineshape_voigt_lte.cpp:1562-1637): one catalog line's f0, e0, Einstein
   * coefficient, or one coefficient X0..X3 of one broadener's G0 / D0 / DV / Y / G model (G2, D2, FVC, ETA give zero
   * rows, :1616-1619).  Only sub-lines of that line contribute (set_filter :1192-1201).  Bands with a ByLine cutoff are
   * AB200_ERR_UNSUPPORTED: the reference indexes the cutoff window's sub-span with whole-band indices there (:723-739) */
  AB200_TARGET_LINE_F0 = 8,
  AB200_TARGET_LINE_E0 = 9,
  AB200_TARGET_LINE_A = 10,
  AB200_TARGET_LINE_LS = 11,
  /* AtmKey::p: "Not implemented, pressure derivative" in the reference (:1482); AB200_ERR_UNSUPPORTED with that text */
  AB200_TARGET_P = 12,
  /* SpeciesIsotope (isotopologue ratio): scl shape / ratio of the bands of that isotopologue (:1526-1544);
   * .species holds the ISOTOPOLOGUE index; a zero ratio is the reference's error */
  AB200_TARGET_ISORAT = 13
};

/* flags (bit mask) */
#define AB200_FLAG_K_ZERO_INIT 1u /* caller's K/dK are known to be zero: skip their H2D (+= still holds) */
#define AB200_FLAG_TRAN_EXACT 2u  /* use the exact Cayley-Hamilton eigen pair instead of the reference's \
                                     literal rtepack_transmission.cc:64-70 arithmetic */
#define AB200_FLAG_RETURN_K 4u    /* clearsky_emission: also copy K back to the host */
#define AB200_FLAG_NO_EMISSION 8u /* fused path: pure transmission, J = 0 at every level (spectral_radCumulativeTransmission) */
#define AB200_FLAG_WIND_ROWS_DF 16u /* wind rows stay the frequency derivative d propmat / d f (what                   \
                                      spectral_propmatAddLines leaves in spectral_propmat_jac): for a caller whose agenda   \
                                      runs spectral_propmat_jacWindFix itself.  Without it the rows are d propmat / d wind */

/* ---- catalog: AbsorptionBands flattened (lbl_data.h:31-68,196-300) ----- */
typedef struct ab200_catalog_desc {
  int32_t n_species; /* size of the vmr vector per level; species ids are 0..n_species-1 */
  int32_t n_isot;    /* isotopologues referenced by bands */
  int32_t n_bands;
  int64_t n_lines;
  int64_t n_ls; /* total number of (line, broadener) entries */

  const int32_t *isot_species; /* [n_isot] species id of each isotopologue */
  const double *isot_mass;     /* [n_isot] g/mol (SpeciesIsotope::mass) */

  const int32_t *band_isot;         /* [n_bands] */
  const int32_t *band_lineshape;    /* [n_bands] AB200_LINESHAPE_* */
  const int32_t *band_cutoff_type;  /* [n_bands] AB200_CUTOFF_* */
  const double *band_cutoff_value;  /* [n_bands] Hz */
  const int64_t *band_offset;       /* [n_bands+1] lines of band b are [band_offset[b], band_offset[b+1]) */

  const double *f0, *a, *e0, *gu, *gl; /* [n_lines] lbl::line members */
  const double *T0;                    /* [n_lines] line_shape::model::T0 */

  const uint8_t *z_on;   /* [n_lines] zeeman::model::on */
  const double *z_gu;    /* [n_lines] zeeman g upper */
  const double *z_gl;    /* [n_lines] zeeman g lower */
  const int32_t *two_Ju; /* [n_lines] 2*J upper (Rational) */
  const int32_t *two_Jl; /* [n_lines] 2*J lower */

  const int64_t *ls_offset;  /* [n_lines+1] broadeners of line l are [ls_offset[l], ls_offset[l+1]) */
  const int32_t *ls_species; /* [n_ls] species id or AB200_SPECIES_BATH */
  const int32_t *ls_type;    /* [n_ls][AB200_NVAR] AB200_TM_* (ABSENT if the variable is not in the map) */
  const double *ls_X;        /* [n_ls][AB200_NVAR][4] X0..X3 */
} ab200_catalog_desc;

/* ---- ArrayOfAtmPoint + ArrayOfPropagationPathPoint flattened ------------ */
typedef struct ab200_atm_path {
  int32_t np;
  const double *T;      /* [np] K */
  const double *P;      /* [np] Pa */
  const double *vmr;    /* [np][n_species] */
  const double *isorat; /* [np][n_isot] isotopologue ratios atm[spec] */
  const double *Q;      /* [np][n_isot] PartitionFunctions::Q(T, isot) */
  const double *dQdT;   /* [np][n_isot] PartitionFunctions::dQdT(T, isot); may be NULL if no T target */
  const double *mag;    /* [np][3] magnetic field u,v,w [T]; may be NULL (=0) */
  const double *los;    /* [np][2] zenith, azimuth [deg] (PropagationPathPoint::los); may be NULL (=0) */
  const double *wind;   /* [np][3] wind u,v,w [m/s] (AtmPoint::wind); may be NULL (=0).  With winds the library applies
                           freq_grid_pathFromPath (src/m_ppvar.cc:47-77, wind_shift src/m_frequency_grid.cc:4-84) on the
                           device: level ip sees fac[ip] * f with fac = 1 - |wind| cos(angle(wind, mirrored los)) / c,
                           so the caller passes ONE grid instead of np shifted copies */
} ab200_atm_path;

typedef struct ab200_target {
  int32_t kind;    /* AB200_TARGET_* */
  int32_t species; /* AB200_TARGET_VMR: the species; AB200_TARGET_LINE_LS: the broadener (or AB200_SPECIES_BATH);
                      AB200_TARGET_ISORAT: the isotopologue */
  int64_t line;    /* AB200_TARGET_LINE_*: index of the line in the flattened catalog (band_offset order) */
  int32_t ls_var;  /* AB200_TARGET_LINE_LS: AB200_VAR_* */
  int32_t coeff;   /* AB200_TARGET_LINE_LS: 0..3 for X0..X3 (LineShapeModelCoefficient) */
} ab200_target;

typedef struct ab200_catalog ab200_catalog; /* opaque, immutable after create, shareable across threads */
typedef struct ab200_path ab200_path;       /* opaque device-resident path workspace */

/* message of the last failure on the calling thread */
const char *ab200_last_error(void);

/* number of usable CUDA devices (0 -> every compute call fails) */
int ab200_device_count(void);
/* Makes `device` current for the calling host thread (one process per GPU: call once with LOCAL_RANK).
 * Catalogs and path workspaces are created on the device that is current at the time. */
int ab200_set_device(int device);

/* Copies the description to the current CUDA device as SoA + pre-expanded
 * Zeeman sub-lines (lbl_zeeman.cpp:261-309, lbl_zeeman.h:342-352).
 * Replaces: the AoS-of-maps walk in band_shape_helper (lbl_lineshape_voigt_lte.cpp:394-429). */
int ab200_catalog_create(const ab200_catalog_desc *desc, ab200_catalog **out);
void ab200_catalog_destroy(ab200_catalog *cat);
/* number of (sub-)lines after Zeeman expansion, per polarisation no,pi,sm,sp */
int ab200_catalog_counts(const ab200_catalog *cat, int64_t counts[4]);

/* spectral_propmatAddLines (np == 1) and spectral_propmat_pathFromPath with
 * the lines-only agenda (all levels in one call).
 *   f:  [np][nf] when f_level_stride == nf (freq_grid_path), or [nf] shared when 0
 *   K:  [np][nf][7]  accumulated (+=) like lbl_lineshape_voigt_lte.cpp:1688-1692
 *   dK: [np][nq][nf][7] accumulated (+=) (PropmatMatrix [nq,nf] per level); NULL if nq == 0
 * all host pointers. */
int ab200_propmat_levels(const ab200_catalog *cat, int64_t nf, const double *f, int64_t f_level_stride,
                         const ab200_atm_path *atm, int32_t select_species, int32_t no_negative_absorption,
                         int32_t nq, const ab200_target *targets, uint32_t flags, double *K, double *dK);

/* spectral_tramat_pathFromPath -> TransmittanceMatrix::init (rtepack_transmission.cc:1254-1328).
 *   K [np][nf][7], dK [np][nq][nf][7], r [np-1], dr [2][np-1][nq]
 *   T,L,P [nf][np][16] (index 0 = identity), dT,dL [2][nf][np][nq][16]; L/dL may be NULL for constant. */
int ab200_tramat(int32_t np, int64_t nf, int32_t nq, const double *K, const double *dK, const double *r,
                 const double *dr, int32_t rte_option, uint32_t flags, double *T, double *L, double *P,
                 double *dT, double *dL);

/* spectral_rad_srcvec_pathFromPropmat -> SourceVector::init (rtepack_source.cc:52-105), LTE (S_nlte = 0).
 *   f [np][nf] or [nf] (stride 0), T_level [np], it = index of the T target or -1
 *   J [nf][np][4], dJ [nf][np][nq][4] */
int ab200_srcvec(int32_t np, int64_t nf, int32_t nq, const double *K, const double *f, int64_t f_level_stride,
                 const double *T_level, int32_t it, double *J, double *dJ);

/* spectral_radStepByStepEmission -> rte_emission (rtepack_rtestep.cc:265-404).
 *   I_bkg [nf][4]; I [nf][4]; dI [nf][np][nq][4] */
int ab200_rte_emission(int32_t rte_option, int32_t np, int64_t nf, int32_t nq, const double *T, const double *L,
                       const double *P, const double *dT, const double *dL, const double *J, const double *dJ,
                       const double *I_bkg, double *I, double *dI);

/* The fused fast path: the canonical sequence of spectral_radClearskyEmission
 * (workspace_meta_methods.cpp:166-181) from the propagation matrix to the
 * radiance.  K, T, Lambda, J never leave the device / registers.
 *   r [np-1] layer lengths, hse_derivative as in m_tramat.cc:18-24
 *   I [nf][4], dI [nf][np][nq][4] (NULL if nq==0), K_out [np][nf][7] only with AB200_FLAG_RETURN_K */
int ab200_clearsky_emission(const ab200_catalog *cat, int64_t nf, const double *f, int64_t f_level_stride,
                            const ab200_atm_path *atm, int32_t select_species, int32_t no_negative_absorption,
                            int32_t nq, const ab200_target *targets, const double *r, int32_t hse_derivative,
                            int32_t rte_option, const double *I_bkg, uint32_t flags, double *I, double *dI,
                            double *K_out);

/* spectral_radCumulativeTransmission (src/m_spectral_radiance.cc:49-74 -> rte_transmission,
 * rtepack_rtestep.cc:456-503): spectral_rad = P[.,np-1] * spectral_rad_bkg and the transmission-only
 * spectral_rad_jac_path.  T, P [nf][np][16], dT [2][nf][np][nq][16] as produced by ab200_tramat.
 * The Jacobian is the derivative of the transmitted radiance (the emission recursion with J = 0); the reference's
 * loop indexes the layer transmittance as Ts[i+1][iv] (frequency and level swapped) and files the dT[1] term one
 * level early, so for nq > 0 there is no defined reference output to be equal to. */
int ab200_rte_transmission(int32_t np, int64_t nf, int32_t nq, const double *T, const double *P, const double *dT,
                           const double *I_bkg, double *I, double *dI);

/* spectral_radApplyUnitFromSpectralRadiance with PlanckBT
 * (spectral_radiance_transform_operator.cc:46-87): in place on I [nf][4]. */
int ab200_planck_tb(int64_t nf, const double *f, double *I);

/* ---- device-resident path workspace (what clearsky_emission is built on) --- */
/* stream: a cudaStream_t (as void*) the caller wants the kernels on, or NULL for the library's own. */
int ab200_path_create(const ab200_catalog *cat, int64_t nf, int32_t np, int32_t nq, ab200_path **out);
/* A workspace that only runs the Stokes chain (and downloads): no line records, cluster moments or Jacobian scratch.  Its K
 * [np][k_pitch][7], k_pitch = nf rounded up to 128 (ab200_path_device_ptr(p, 1)), is filled by the caller with rows that other
 * workspaces of the SAME catalog, species selection and flags summed - e.g. after an exchange between devices that deal the
 * levels of the line sum - and handed over with ab200_path_adopt_K after ab200_path_upload. */
int ab200_path_create_stage2(const ab200_catalog *cat, int64_t nf, int32_t np, ab200_path **out);
int ab200_path_adopt_K(ab200_path *p);
void ab200_path_destroy(ab200_path *p);
int ab200_path_set_stream(ab200_path *p, void *stream);
/* H2D of one path's inputs (asynchronous on the path's stream; small arrays go through pinned staging owned by the
 * workspace, f and I_bkg are copied straight from the caller's buffers, which must stay unchanged until the next
 * ab200_path_sync / ab200_path_download if they are pinned). */
int ab200_path_upload(ab200_path *p, const double *f, int64_t f_level_stride, const ab200_atm_path *atm,
                      int32_t select_species, int32_t no_negative_absorption, const ab200_target *targets,
                      const double *r, int32_t hse_derivative, int32_t rte_option, const double *I_bkg,
                      uint32_t flags);
/* Frequency sharding: the ByLine line selection band_data::active_lines (lbl_data.cpp:61-68) uses the first
 * and last frequency of the grid of the call (lbl_lineshape_voigt_lte.cpp:1672-1680).  A rank that uploads
 * only its shard of the grid passes the bounds of the WHOLE grid here (before ab200_path_upload) so that the
 * result is bit-identical for every shard count.  bounds: [np][2] = {f_first, f_last} per level, or NULL to
 * go back to the bounds of the uploaded grid. */
int ab200_path_set_grid_bounds(ab200_path *p, const double *bounds);
/* launches K1 (prepare) + K2/K3 (line sum) into the resident K, asynchronously */
int ab200_path_run_propmat(ab200_path *p);
/* launches the fused K4-K7 Stokes chain on the resident K, asynchronously */
int ab200_path_run_stokes(ab200_path *p);
/* D2H of the results (synchronises the stream). Any pointer may be NULL. */
int ab200_path_download(ab200_path *p, double *I, double *dI, double *K, double *dK);
int ab200_path_sync(ab200_path *p);
/* raw device pointers (for NCCL gathers by the caller): which = 0: I [nf][4], 1: K [np][nf][7], 2: dI, 3: dK */
void *ab200_path_device_ptr(ab200_path *p, int which);
/* kernels launched by the library on this thread since the last call (bench.py's gpu_launches) */
int64_t ab200_launch_count(int reset);

/* ---- collision-induced absorption: spectral_propmatAddCIA on the resident K -----------------------
 * src/m_cia.cc:27-178 with CIARecord::Extract (src/core/absorption/cia.cc:214-226) and cia_interpolation (:76-190):
 * every dataset of a species pair is a GriddedField2 [frequency, temperature]; third-order Lagrange interpolation in
 * frequency (extrapolation factor 0.5), order min(3, nT - 1) in temperature (T_extrapolfac), zero outside the dataset's
 * frequency range, negative overshoots clamped to zero, datasets added;
 *   K.A += xsec * nd^2 * vmr[species1] * vmr[species2],   nd = P / (k T)
 * with the reference's temperature Jacobian by perturbation (xsec(T + dT) - xsec(T)) / dT and its VMR Jacobians.
 * Interpolation: lagrange_interp::make_lags / interp (src/core/matpack/lagrange_interp.h:160-248,300-440,572-650,920-940). */
typedef struct ab200_cia_dataset {
  int32_t nf, nT;
  const double *f_grid; /* [nf] ascending, nf >= 4 */
  const double *T_grid; /* [nT] ascending */
  const double *data;   /* [nf][nT] binary absorption cross-section */
} ab200_cia_dataset;
typedef struct ab200_cia_record {
  int32_t species1, species2; /* SpeciesEnumPair, the caller's species indices */
  int32_t n_datasets;
  const ab200_cia_dataset *datasets;
} ab200_cia_record;
typedef struct ab200_cia ab200_cia; /* device copy, immutable, shareable */
int ab200_cia_create(const ab200_cia_record *records, int32_t n_records, ab200_cia **out);
void ab200_cia_destroy(ab200_cia *cia);
/* Adds the CIA term to the resident K (and to dK of the path's temperature / VMR targets) of an uploaded path, after
 * ab200_path_run_propmat.  dT: perturbation of the temperature target (JacobianTargets' `d`), ignored without one.
 * A temperature outside a dataset's extrapolation range is the reference's error unless ignore_errors (then NaN). */
int ab200_path_add_cia(ab200_path *p, const ab200_cia *cia, double T_extrapolfac, int32_t ignore_errors, double dT);
/* Host-buffer form (the WSM shim): K [np][nf][7] and dK [np][nq][nf][7] are accumulated (+=) like m_cia.cc:146-177. */
int ab200_cia_levels(const ab200_cia *cia, int64_t nf, const double *f, int64_t f_level_stride, const ab200_atm_path *atm,
                     int32_t n_species, int32_t select_species, int32_t nq, const ab200_target *targets, double dT,
                     double T_extrapolfac, int32_t ignore_errors, double *K, double *dK);

/* ---- absorption lookup tables: spectral_propmatAddLookup on the resident K -----------------------
 * lookup::table (src/core/lookup/lookup_map.{h,cpp}): per species a cross-section tensor xsec [t_pert][w_pert][p][f]
 * precomputed with the line-by-line code on a reference profile (table ctor :22-131 = K.A / number density, what
 * ab200_propmat_levels delivers for the perturbed profiles), extracted by Lagrange interpolation of the given orders
 * in temperature offset, water ratio, log-pressure (descending grid) and frequency (table::absorption :190-238 with
 * pressure_/frequency_/water_/temperature_lagrange :133-188), times the species' number density.
 * _spectral_propmatAddLookup (src/m_lookup.cc:20-141): K.A += absorption where no_negative_absorption == 0 or it is
 * positive; every Jacobian target by re-extraction at the perturbed point, ASSIGNED to the row (sic, :130-135). */
struct ab200_partfun_table;
typedef struct ab200_lookup_table {
  int32_t species;           /* the species this table stands for (atm_point.number_density(species)) */
  int32_t nf, np, nt, nw;    /* nt / nw: size of t_pert / w_pert, or 1 when the table has no such grid (do_t / do_w false) */
  int32_t do_t, do_w;
  const double *f_grid;      /* [nf] ascending */
  const double *log_p_grid;  /* [np] descending */
  const double *t_pert;      /* [nt] ascending (do_t) */
  const double *w_pert;      /* [nw] ascending (do_w) */
  const double *t_atmref;    /* [np] temperature of the reference profile */
  const double *water_atmref;/* [np] H2O VMR of the reference profile (do_w) */
  const double *xsec;        /* [nt][nw][np][nf] */
} ab200_lookup_table;
typedef struct ab200_lookup ab200_lookup; /* device copy of AbsorptionLookupTables */
int ab200_lookup_create(const ab200_lookup_table *tables, int32_t n_tables, ab200_lookup **out);
void ab200_lookup_destroy(ab200_lookup *lut);
/* Host-buffer form.  h2o_species: index of H2O in the vmr vector (-1 if no table has a water grid).  target_d [nq]: the
 * perturbation of every Jacobian target (JacobianTargets' `d`).  K [np][nf][7] +=, dK [np][nq][nf][7] rows assigned. */
int ab200_lookup_levels(const ab200_lookup *lut, int64_t nf, const double *f, int64_t f_level_stride,
                        const ab200_atm_path *atm, int32_t n_species, int32_t h2o_species, int32_t select_species,
                        int32_t nq, const ab200_target *targets, const double *target_d, int32_t no_negative_absorption,
                        int32_t p_interp_order, int32_t t_interp_order, int32_t water_interp_order, int32_t f_interp_order,
                        double extpolfac, double *K, double *dK);

/* abs_lookup_dataPrecompute (src/m_lookup.cc:175-197; the lookup::table constructor src/core/lookup/lookup_map.cpp:22-131) with
 * the line-by-line sum on the device: for every temperature offset t_pert[it] (NULL: none, nt = 1) and water ratio w_pert[iw]
 * (NULL: none, nw = 1) the reference profile atm_ref (descending pressure) is perturbed, lbl::calculate runs for select_species
 * with no_negative_absorption = true, and xsec[it][iw][ip][f] = K.A / atm_point.number_density(select_species).  The division
 * runs on the device and only the compact A component comes back: 8 B per table element instead of 56. */
int ab200_lookup_precompute(const ab200_catalog *cat, int64_t nf, const double *f, const ab200_atm_path *atm_ref,
                            int32_t select_species, int32_t h2o_species, int32_t nt, const double *t_pert, int32_t nw,
                            const double *w_pert, const struct ab200_partfun_table *partfun, double *xsec);
/* partfun: [n_isot] partition-function tables (below): the reference evaluates Q at the PERTURBED temperature inside
 * lbl::calculate, so Q(T + t_pert) is formed here with ab200_partfun_eval.  NULL: atm_ref->Q is used for every offset
 * (exact only without a temperature grid). */

/* On the resident path: adds the lookup-table absorption to K (and assigns the rows of dK of the path's targets).  With
 * zero_init the resident K / dK are cleared first - the agenda with use_abs_lookup_data = 1 has no line-by-line term
 * (src/m_abs.cc:257-266), so ab200_path_run_propmat is simply not called.  no_negative_absorption is the path's. */
int ab200_path_add_lookup(ab200_path *p, const ab200_lookup *lut, int32_t h2o_species, const double *target_d,
                          int32_t p_interp_order, int32_t t_interp_order, int32_t water_interp_order,
                          int32_t f_interp_order, double extpolfac, int32_t zero_init);

/* ---- predefined continua: spectral_propmatAddPredefined on the resident K -------------------------
 * src/m_predefined_absorption_models.cc:156-191 -> Absorption::PredefinedModel::compute
 * (src/core/absorption/predefined_absorption_models.cc:219-317): the model's closed form added to K.A, the temperature row
 * by perturbation (model(T + d) - model(T)) / d and VMR rows by perturbation for targets of CO2, O2, N2, H2O and
 * liquidcloud only (:237-241, compute_vmr_deriv :202-216).  Models on the path: the four "StandardType" continua of
 * src/core/predefined/standard.cc (Rosenkranz 1993 / 1998) and the full microwave models PWR98 (H2O, O2), MPM89 (H2O, O2),
 * the MPM93 N2 continuum and Rosenkranz's 2021 / 2022 revisions (H2O, O2, N2); every other model name is AB200_ERR_UNSUPPORTED. */
#define AB200_PREDEF_O2_SELFCONT_STANDARD 0     /* "O2-SelfContStandardType",     Standard::oxygen        standard.cc:51-84 */
#define AB200_PREDEF_N2_SELFCONT_STANDARD 1     /* "N2-SelfContStandardType",     Standard::nitrogen      :118-138 */
#define AB200_PREDEF_H2O_FOREIGNCONT_STANDARD 2 /* "H2O-ForeignContStandardType", Standard::water_foreign :166-184 */
#define AB200_PREDEF_H2O_SELFCONT_STANDARD 3    /* "H2O-SelfContStandardType",    Standard::water_self    :212-226 */
#define AB200_PREDEF_H2O_PWR98 4          /* "H2O-PWR98",        PWR98::water    src/core/predefined/PWR98.cc:40-242  (15 lines + continuum) */
#define AB200_PREDEF_O2_PWR98 5           /* "O2-PWR98",         PWR98::oxygen   PWR98.cc:297-434 (40 lines with mixing + dry continuum) */
#define AB200_PREDEF_H2O_MPM89 6          /* "H2O-MPM89",        MPM89::water    MPM89.cc:95-180  (30 lines + continuum) */
#define AB200_PREDEF_O2_MPM89 7           /* "O2-MPM89",         MPM89::oxygen   MPM89.cc:270-411 (44 lines with mixing + Debye term) */
#define AB200_PREDEF_N2_SELFCONT_MPM93 8  /* "N2-SelfContMPM93", MPM93::nitrogen MPM93.cc:33-73 */
#define AB200_PREDEF_H2O_PWR2021 9           /* "H2O-PWR2021",        PWR20xx::compute_h2o_2021 src/core/predefined/PWR20xx.cc:169-381 (16 lines) */
#define AB200_PREDEF_H2O_PWR2022 10          /* "H2O-PWR2022",        PWR20xx::compute_h2o_2022 :383-491 (20 lines; shape :21-166) */
#define AB200_PREDEF_O2_PWR2021 11           /* "O2-PWR2021",         PWR20xx::compute_o2_2021  :576-682 (49 lines; shape :494-573) */
#define AB200_PREDEF_O2_PWR2022 12           /* "O2-PWR2022",         PWR20xx::compute_o2_2022  :684-790 */
#define AB200_PREDEF_N2_SELFCONT_PWR2021 13  /* "N2-SelfContPWR2021", PWR20xx::compute_n2       :792-833 */
#define AB200_PREDEF_O2_TRE05 14             /* "O2-TRE05",           TRE05::oxygen src/core/predefined/TRE05.cc:115-296 (44 lines, MPM93 form) */
#define AB200_PREDEF_O2_MPM2020 15           /* "O2-MPM2020",         MPM2020::compute src/core/predefined/MPM2020.cc:38-149 (38 lines, second-order mixing) */
#define AB200_PREDEF_LIQUIDCLOUD_ELL07 16    /* "liquidcloud-ELL07",  ELL07::compute src/core/predefined/ELL07.cc:39-188: Ellison (2007) permittivity of liquid
                                                water (three Debye terms + two resonances), Rayleigh droplets; the "mixing ratio" of the species
                                                `liquidcloud` is the liquid water content [kg/m3].  Nothing below 1e-10 kg/m3; above 5e-3 kg/m3, outside
                                                210-373 K or above 25 THz the reference's user error (AB200_ERR_INVALID) */
#define AB200_PREDEF_H2O_FOREIGNCONT_CKDMT400 17 /* "H2O-ForeignContCKDMT400", MT_CKD400::compute_foreign_h2o src/core/predefined/MT_CKD400.cc:99-172 */
#define AB200_PREDEF_H2O_SELFCONT_CKDMT400 18    /* "H2O-SelfContCKDMT400",    MT_CKD400::compute_self_h2o    :174-256 */
#define AB200_PREDEF_H2O_FOREIGNCONT_CKDMT430 19 /* "H2O-ForeignContCKDMT430", MT_CKD430::compute_foreign_h2o src/core/predefined/MT_CKD430.cc (same arithmetic) */
#define AB200_PREDEF_H2O_SELFCONT_CKDMT430 20    /* "H2O-SelfContCKDMT430",    MT_CKD430::compute_self_h2o */
typedef struct ab200_predef_species { /* indices into the vmr vector, -1 when the atmosphere does not carry the species */
  int32_t o2, n2, h2o, co2, liquidcloud;
} ab200_predef_species;
/* models [n_models]: AB200_PREDEF_*; target_d [nq]: perturbation of every Jacobian target.  K [np][nf][7], dK [np][nq][nf][7] +=. */
int ab200_predef_levels(const int32_t *models, int32_t n_models, const ab200_predef_species *species, int64_t nf,
                        const double *f, int64_t f_level_stride, const ab200_atm_path *atm, int32_t n_species,
                        int32_t select_species, int32_t nq, const ab200_target *targets, const double *target_d,
                        double *K, double *dK);
int ab200_path_add_predefined(ab200_path *p, const int32_t *models, int32_t n_models, const ab200_predef_species *species,
                              const double *target_d);

/* Model data of the MT_CKD 4.x water continua (PredefinedModelData, src/core/predefined/predef_data.h:14-42): the coefficient
 * tables the user loads with abs_predef_dataAddWaterMTCKD400 / ...430 (src/m_predefined_absorption_models.cc:69-148; same
 * checks: equal lengths >= 4, ascending wavenumbers) on a regular wavenumber grid [cm-1].  The continuum of a frequency is the
 * four-point interpolation XINT_FUN (MT_CKD400.cc:84-92) of the scaled coefficients around its wavenumber, times the radiation
 * term RADFN_FUN (:37-77); the first table entry is mirrored below the grid and frequencies beyond the last wavenumber get
 * nothing (:139-171).  The reference walks the table with a cursor along the ascending grid; here every frequency finds its
 * own interval (same interval, same four coefficients). */
typedef struct ab200_mtckd_water {
  int32_t n;                     /* table length */
  double ref_temp, ref_press;    /* [K], [hPa] */
  const double *wavenumbers;     /* [n] cm-1, regular */
  const double *self_absco_ref;  /* [n] */
  const double *for_absco_ref;   /* [n] */
  const double *self_texp;       /* [n] */
} ab200_mtckd_water;
typedef struct ab200_predef_data ab200_predef_data; /* device-resident copy of the tables */
/* ckdmt400 / ckdmt430: the data of the "...CKDMT400" / "...CKDMT430" tags, NULL when not loaded (a model without its data is
 * the reference's "No data" error).  MT_CKD430::WaterData also carries for_closure_absco_ref, which none of the dispatched
 * functions reads (predefined_absorption_models.cc:62-77); it only has to be non-empty there, so it is not passed. */
int ab200_predef_data_create(const ab200_mtckd_water *ckdmt400, const ab200_mtckd_water *ckdmt430, int32_t device,
                             ab200_predef_data **out);
void ab200_predef_data_destroy(ab200_predef_data *data);
/* ab200_predef_levels / ab200_path_add_predefined with model data (data may be NULL: same as the calls above) */
int ab200_predef_levels_data(const int32_t *models, int32_t n_models, const ab200_predef_species *species, int64_t nf,
                             const double *f, int64_t f_level_stride, const ab200_atm_path *atm, int32_t n_species,
                             int32_t select_species, int32_t nq, const ab200_target *targets, const double *target_d,
                             double *K, double *dK, const ab200_predef_data *data);
int ab200_path_add_predefined_data(ab200_path *p, const int32_t *models, int32_t n_models, const ab200_predef_species *species,
                                   const double *target_d, const ab200_predef_data *data);

/* ---- catalog ingest: HITRAN .par records straight into the SoA of ab200_catalog_desc -------------
 * abs_bandsReadHITRAN (src/m_lbl.cc:302-338) with file_formatter = ["par"], either line_strength_option,
 * compute_zeeman_parameters = 0: read_par_line (src/core/lbl/lbl_hitran.cpp:66-89, the 160-column record and its unit
 * conversions), read_hitran_par (:146-172: records below frequency_range[0] are skipped, reading stops at the first
 * one above frequency_range[1]) and hitran_record::from (:180-237: T0 = 296 K, G0 = T1(gamma, n) for the line's own
 * species and for Bath (air), D0 = T0(delta) for both when delta != 0).  Without quantum-number columns the reference
 * keys its bands by isotopologue: one band per isotopologue, in the order of the table, lines in file order.
 * The records are parsed by n_threads host threads (0: all cores) without building the reference's AoS of maps. */
typedef struct ab200_hitran_isotopologue {
  int32_t M;       /* HITRAN molecule number (columns 1-2) */
  char I;          /* HITRAN isotopologue character (column 3) */
  int32_t species; /* the caller's species index of this isotopologue (Hitran::id_from_lookup + SpeciesEnum) */
  double mass;     /* g/mol */
  double hitran_ratio; /* Hitran::isotopologue_ratios()[isot]; used by AB200_HITRAN_STRENGTH_S only */
  double Q296;         /* PartitionFunctions::Q(296, isot);    used by AB200_HITRAN_STRENGTH_S only */
} ab200_hitran_isotopologue;
/* HitranLineStrengthOption: S (the reference's default) turns the record's line strength into the Einstein coefficient,
 * a = einstein_a(S / ratio, gu, e0, f0, 296 K, Q(296)) (line::hitran_a lbl_data.cpp:155-169, einstein_a :34-40;
 * g_upp == 0 becomes gu = gl = -1, lbl_hitran.cpp:193-200); A takes the file's own coefficient */
enum { AB200_HITRAN_STRENGTH_S = 0, AB200_HITRAN_STRENGTH_A = 1 };
typedef struct ab200_hitran_catalog ab200_hitran_catalog; /* owns the arrays the description points to */
int ab200_hitran_read_par(const char *text, int64_t len, double fmin, double fmax, int32_t line_strength_option,
                          const ab200_hitran_isotopologue *isotopologues, int32_t n_isot, int32_t n_species,
                          int32_t n_threads, ab200_hitran_catalog **out);
int ab200_hitran_read_par_file(const char *filename, double fmin, double fmax, int32_t line_strength_option,
                               const ab200_hitran_isotopologue *isotopologues, int32_t n_isot, int32_t n_species,
                               int32_t n_threads, ab200_hitran_catalog **out);
/* the description to hand to ab200_catalog_create (valid until ab200_hitran_destroy) */
const ab200_catalog_desc *ab200_hitran_desc(const ab200_hitran_catalog *cat);
void ab200_hitran_destroy(ab200_hitran_catalog *cat);

/* ---- catalog ingest: the reference's own AbsorptionBands XML straight into the SoA ----------------
 * xml_io_stream<AbsorptionBand>::read (src/core/lbl/lbl_data.cpp:435-470) inside the
 * <Map type="AbsorptionBand" key="QuantumIdentifier"> of an abs_bands file, every line with operator>>(line)
 * (lbl_data.cpp:52-58: f0 a e0 gu gl, zeeman::model lbl_zeeman.cpp:311-319, line_shape::model
 * lbl_lineshape_model.cpp:260-296 with temperature::data lbl_temperature_model.cpp:28-43, local quantum numbers
 * quantum.cc:150-163).  One band per <AbsorptionBand> in file order.  Names are resolved through the caller's tables:
 * isotopologue tags ("H2O-161") and broadener names in whatever spelling the files use ("Nitrogen", "N2", "Bath").
 * G2 / D2 / FVC / ETA entries must be all-zero (dropped, like model::clear_zeroes), POLY takes at most four coefficients,
 * a line with Zeeman on needs its local J; anything else the GPU path cannot hold is AB200_ERR_UNSUPPORTED. */
typedef struct ab200_xml_isotopologue {
  const char *name; /* SpeciesIsotope tag, e.g. "O2-66" */
  int32_t species;  /* the caller's species index */
  double mass;      /* g/mol */
} ab200_xml_isotopologue;
typedef struct ab200_xml_species {
  const char *name; /* SpeciesEnum name as written in the file */
  int32_t species;  /* species index or AB200_SPECIES_BATH */
} ab200_xml_species;
typedef struct ab200_xml_catalog ab200_xml_catalog; /* owns the arrays the description points to */
int ab200_xml_read_bands(const char *text, int64_t len, const ab200_xml_isotopologue *isotopologues, int32_t n_isot,
                         const ab200_xml_species *names, int32_t n_names, int32_t n_species, ab200_xml_catalog **out);
int ab200_xml_read_bands_file(const char *filename, const ab200_xml_isotopologue *isotopologues, int32_t n_isot,
                              const ab200_xml_species *names, int32_t n_names, int32_t n_species, ab200_xml_catalog **out);
const ab200_catalog_desc *ab200_xml_desc(const ab200_xml_catalog *cat);
void ab200_xml_destroy(ab200_xml_catalog *cat);

/* ---- partition functions: Q(T) and dQ/dT of every isotopologue at every level ---------------------
 * PartitionFunctions::Q / dQdT (src/partfun/partfun.h) are generated at build time from data tables by
 * src/partfun/make_auto_partfuns.cc; the four table kinds and their literal formulas:
 *   INTERP        :28-63   linear interpolation on an increasing grid, i = min(lower_bound(T) - (>0), n - 2)
 *   COEFF         :65-105  polynomial sum_i Q[i] T^i, derivative sum_i i Q[i] T^(i-1) (running power, same order)
 *   CONST         :107-115 a constant, derivative 0
 *   STATIC_INTERP :117-153 equidistant grid: Tx = (T - T0) * (1 / dT), i = min(size_t(Tx), n - 2)
 * Fills ab200_atm_path.Q and .dQdT from the tables so that the shim need not evaluate them per level. */
#define AB200_PARTFUN_INTERP 0
#define AB200_PARTFUN_COEFF 1
#define AB200_PARTFUN_CONST 2
#define AB200_PARTFUN_STATIC_INTERP 3
typedef struct ab200_partfun_table {
  int32_t kind;       /* AB200_PARTFUN_* */
  int32_t n;          /* number of grid points (INTERP, STATIC_INTERP) or coefficients (COEFF); 1 for CONST */
  const double *grid; /* [n] temperatures (INTERP, STATIC_INTERP), else NULL */
  const double *coef; /* [n] Q values, polynomial coefficients, or the constant */
} ab200_partfun_table;
/* Q, dQdT: [np][n_isot] like in ab200_atm_path; dQdT may be NULL. */
int ab200_partfun_eval(const ab200_partfun_table *tables, int32_t n_isot, int32_t np, const double *T, double *Q, double *dQdT);

/* ---- atm_pathFromPath for a 1-D atmosphere (src/m_ppvar.cc:38-45) ------------------------------------
 * forward_atm_path (src/core/path/atm_path.cpp:19-28): AtmField::at at every path point, with the top of the atmosphere for
 * points outside it.  The AtmField of a 1-D atmosphere - every key a GeodeticField3 on one altitude grid with 1 x 1
 * latitude / longitude, isotopologue ratios plain numbers - is flattened ONCE into this struct; a path then costs np linear
 * (Lagrange order 1, functional_atm_field_interp.cpp:6-10) interpolations on flat arrays instead of a walk over the
 * AtmPoint maps per path.  Out-of-grid altitudes follow the field's InterpolationExtrapolation (atm_field.cpp:536-566,
 * :890-924): None is an error, Zero gives 0, Nearest the boundary value, Linear the end stencil; an altitude above
 * top_of_atmosphere is an error (:928-935), as are NaN temperature / pressure and NaN or negative VMRs (:601-640).
 * Host function (no GPU needed): a path is O(np) numbers and ab200_path_upload consumes host arrays. */
#define AB200_EXTRAP_NONE 0
#define AB200_EXTRAP_ZERO 1
#define AB200_EXTRAP_NEAREST 2
#define AB200_EXTRAP_LINEAR 3
typedef struct ab200_atm_profile {
  int32_t nalt;             /* altitude grid points (>= 1; 1: constant, extrapolation Nearest) */
  const double *alt;        /* [nalt] ascending [m] */
  const double *T, *P;      /* [nalt] */
  const double *vmr;        /* [nalt][n_species] */
  const double *isorat;     /* [n_isot] */
  const double *mag, *wind; /* [nalt][3], or NULL (absent from the field: 0) */
  int32_t alt_low, alt_upp; /* AB200_EXTRAP_* below alt[0] / above alt[nalt-1] */
  double top_of_atmosphere; /* AtmField::top_of_atmosphere [m] */
  const struct ab200_partfun_table *partfun; /* [n_isot] Q(T) tables, or NULL (Q, dQdT are then not written) */
} ab200_atm_profile;
/* alt [np] = pos[0] of the path points, in_atm [np] = PropagationPathPoint::has(PathPositionType::atm) (NULL: all inside).
 * Outputs are the arrays of ab200_atm_path: T, P [np], vmr [np][n_species], isorat, Q, dQdT [np][n_isot], mag, wind [np][3];
 * dQdT, mag, wind may be NULL. */
int ab200_atm_path_from_profile(const ab200_atm_profile *prof, int32_t n_species, int32_t n_isot, int32_t np, const double *alt,
                                const uint8_t *in_atm, double *T, double *P, double *vmr, double *isorat, double *Q,
                                double *dQdT, double *mag, double *wind);

/* ---- observer epilogue on the device (the callers' glue around the path) --------------------
 * What spectral_rad_observer_agenda / measurement_vecFromSensor do on the host after the RTE, applied to the
 * resident results of one path so that only the state-space Jacobian or the sensor channels cross PCIe:
 *   1. background radiance from a temperature: spectral_radSurfaceBlackbody / spectral_radUniformCosmicBackground
 *      (src/m_background.cc:55-71,113-141), with the surface-temperature rows of spectral_rad_bkg_jac;
 *   2. spectral_rad_jacFromBackground (src/m_rad.cc:26-60): Jx[i][f] = P[f][np-1] * bkg_jac[i][f];
 *   3. spectral_rad_jacAddPathPropagation (src/m_rad.cc:62-127): Jx[i][f] = fma(w, dI[f][ip][t], Jx[i][f]) with the
 *      field's flat interpolation weights of every path point (computed by the shim: AtmField::flat_weight);
 *   4. spectral_rad_transform_operator (spectral_radiance_transform_operator.cc:8-122) on I and Jx;
 *   5. SensorObsel::sumup (src/core/sensor/obsel.cpp:246-279) of this path's poslos row: y[c] and Jy[c][i].
 * The x-space accumulation runs inside the fused Jacobian pass (the per-level dI is never written). */
#define AB200_UNIT_UNIT 0       /* spectral_unit_op :8-19 */
#define AB200_UNIT_RJBT 1       /* spectral_rjbt_op :21-44 */
#define AB200_UNIT_PLANCKBT 2   /* spectral_planck_op :46-87 */
#define AB200_UNIT_W_M2_M_SR 3  /* spectral_W_m2_m_sr_op :89-112 */
#define AB200_UNIT_W_M2_M1_SR 4 /* spectral_W_m2_m1_sr_op :114-122 */
#define AB200_BKG_UPLOADED 0    /* the I_bkg given to ab200_path_upload, no background Jacobian */
#define AB200_BKG_PLANCK 1      /* B(f, bkg_T) e_I formed on the device from the sensor's frequency grid */

typedef struct ab200_observer {
  int32_t bkg_kind; /* AB200_BKG_* */
  double bkg_T;     /* surface temperature (single_value of SurfaceKey::t) or the cosmic background's */
  int32_t nx;       /* JacobianTargets::x_size(): rows of spectral_rad_jac */
  /* step 3: CSR over rows ip * nq + t (level, target): x index (x_start included) and weight of every entry */
  const int64_t *map_offset; /* [np * nq + 1] */
  const int32_t *map_x;
  const double *map_w;
  /* steps 1-2: x rows of the surface temperature target, flat_weights at the ground point; bkg_jac = w dB/dT */
  int32_t n_bkg;
  const int32_t *bkg_x;
  const double *bkg_w;
  /* step 4 */
  int32_t unit;  /* AB200_UNIT_* */
  double n_real; /* ray_path.front().nreal */
  /* step 5: CSR over channels of the sparse weight rows with irow == this path's poslos index (sorted by icol) */
  int32_t n_channels; /* 0: no sensor sum-up */
  const int64_t *w_offset; /* [n_channels + 1] */
  const int64_t *w_freq;   /* icol */
  const double *w_stokes;  /* [nnz][4] */
} ab200_observer;

/* Runs the Stokes chain (with the observer's background), the fused Jacobian pass with x-space accumulation, the unit
 * transform and the sensor sum-up on the resident K / dK (after ab200_path_run_propmat), asynchronously.  The index
 * arrays are copied to the device on the path's stream through pageable memory: they may be freed after the call. */
int ab200_path_run_observer(ab200_path *p, const ab200_observer *obs);
/* D2H of the observer results (synchronises).  Any pointer may be NULL.  I [nf][4] transformed spectral_rad;
 * Jx [nx][nf][4] transformed spectral_rad_jac; y [n_channels] and Jy [n_channels][nx]: this path's contribution,
 * which the shim adds to measurement_vec / measurement_jac (src/m_rad.cc:346-351). */
int ab200_path_download_observer(ab200_path *p, double *I, double *Jx, double *y, double *Jy);

/* Stream the calling thread's host-buffer entry points (propmat_levels, clearsky_emission) run on:
 * a cudaStream_t as void*, or NULL for a private non-blocking stream (the default). */
int ab200_set_thread_stream(void *stream);
/* The host-buffer entry points keep one device workspace per calling host thread (the shims are
 * called repeatedly with identical shapes, src/m_rad.cc:321-343).  Drops the calling thread's. */
int ab200_release_thread_cache(void);

/* The host-buffer calls that follow on this thread carry a SHARD of a frequency grid whose first / last frequency per
 * level are bounds [np][2]: ByLine cutoffs select their lines with the bounds of the whole grid
 * (band_data::active_lines, src/core/lbl/lbl_data.cpp:61-68).  np <= 0 or NULL clears it.  For callers that shard the
 * grid themselves; ab200_multi_* uses it for its workers. */
int ab200_set_thread_grid_bounds(int32_t np, const double *bounds);

/* ---- one process, several GPUs ------------------------------------------------------------------------------
 * The reference is ONE process whose frequency loop is an OpenMP team (src/m_lbl.cc:273-295), so a shim inside it makes
 * one call per path with the whole grid.  An ab200_multi holds one catalog replica and one worker thread (own stream,
 * pinned staging, cached workspace) per device; ab200_multi_clearsky_emission / _propmat_levels take exactly the
 * arguments of the single-device calls, deal the grid's 512-frequency blocks round-robin over the devices and let every
 * device write its blocks into the caller's arrays.  No collective; the results are bit-identical to the one-device
 * calls (the value at a frequency does not depend on the shard it is computed in).
 *   n_devices <= 0: every visible device; devices NULL: 0 .. n_devices-1. */
typedef struct ab200_multi ab200_multi;
int ab200_multi_create(const ab200_catalog_desc *desc, int32_t n_devices, const int32_t *devices, ab200_multi **out);
void ab200_multi_destroy(ab200_multi *m);
int32_t ab200_multi_device_count(const ab200_multi *m);
int ab200_multi_clearsky_emission(ab200_multi *m, int64_t nf, const double *f, int64_t f_level_stride,
                                  const ab200_atm_path *atm, int32_t select_species, int32_t no_negative_absorption,
                                  int32_t nq, const ab200_target *targets, const double *r, int32_t hse_derivative,
                                  int32_t rte_option, const double *I_bkg, uint32_t flags, double *I, double *dI,
                                  double *K_out);
int ab200_multi_propmat_levels(ab200_multi *m, int64_t nf, const double *f, int64_t f_level_stride,
                               const ab200_atm_path *atm, int32_t select_species, int32_t no_negative_absorption,
                               int32_t nq, const ab200_target *targets, uint32_t flags, double *K, double *dK);

/* Host-only helpers of the Zeeman pre-expansion (no GPU needed; used by tests and shims):
 * sub-line strengths / splitting coefficients [Hz/T] of one line and polarisation
 * (pol: 0 no, 1 pi, 2 sigma-, 3 sigma+; lbl_zeeman.cpp:261-309, lbl_zeeman.h:342-352); returns the
 * number of sub-lines.  norm_view: the 7-vector of lbl_zeeman.cpp:413-455. */
int ab200_zeeman_components(int on, double gu, double gl, int two_Ju, int two_Jl, int pol, int64_t cap,
                            double *strength, double *splitting);
int ab200_norm_view(int pol, const double *mag, const double *los, double *npm);

/* ---- measurement helpers ------------------------------------------------ */
/* Per-kernel device timing of a path workspace with CUDA events on the path's stream (never under a
 * profiler).  on != 0 starts recording every launch of run_propmat / run_stokes; ab200_path_get_timings
 * synchronises the stream and returns, per kernel class, the accumulated milliseconds and launch counts
 * since the last call: class 0 = line prepare (K1), 1 = real line sum (K2/K3, mode 0), 2 = complex line sum
 * (mode 1), 3 = fused Stokes chain (K4-K6). */
int ab200_path_set_timing(ab200_path *p, int on);
int ab200_path_get_timings(ab200_path *p, double ms[4], int64_t launches[4]);
/* Histogram of the reference's Faddeeva regions (3rdparty/Faddeeva/Faddeeva.cc:689-741,786,890) over a
 * uniform random sample of the path's (line, frequency, level) evaluations — the weights of the
 * algorithmic FLOP count of SURVEY.md 8(d).  out[0..4] = samples in R1 (x+y > 1e7), R2 (> 4000), R3
 * (continued fraction), R4 (series, x < 10), R5 (x >= 10, tiny y); out[5] = sum of the continued-fraction
 * term count nu over the R3 samples; out[6] = samples outside a ByLine cutoff window (not evaluated by the
 * reference); out[7] = samples drawn.  Needs an uploaded path. */
int ab200_path_region_histogram(ab200_path *p, int64_t samples_per_level, uint64_t seed, double out[8]);
/* Dependency-free DFMA loop on all SMs; returns achieved FP64 TFLOP/s (2 flop per DFMA) and the
 * kernel time.  MEASURED_PEAKS.json has no FP64 number. */
int ab200_measure_dfma_peak(int iters, double *tflops, double *ms);
/* Same loop with the far-wing instruction mix: 7 DFMA + 1 MUFU.RCP64H (reciprocal seed) per chain step; returns
 * the DFMA TFLOP/s the FP64 pipe sustains next to the reciprocal seeds (the practical ceiling of the line sum). */
int ab200_measure_dfma_mix(int iters, double *tflops, double *ms);
/* Register-resident w(z) for tests: evaluates the device Faddeeva at n points (host arrays). */
int ab200_faddeeva_w(int64_t n, const double *zr, const double *zi, double *wr, double *wi);
/* Device Faddeeva::Dawson(z) for complex z (3rdparty/Faddeeva/Faddeeva.cc:461-570), the element-wise function of
 * rtepack::dawson(specmat) in polarised linprop layers (rtepack_transmission.cc:467-474): for tests. */
int ab200_dawson(int64_t n, const double *zr, const double *zi, double *dr, double *di);

#ifdef __cplusplus
}
#endif
#endif /* ARTS_B200_H */
