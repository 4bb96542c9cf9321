/* oracle/refslice/wigxjpf_cfg/wigxjpf_auto_config.h - TEST INFRASTRUCTURE ONLY.
 *
 * The one generated header the reference's vendored wigxjpf (3rdparty/wigner/wigxjpf, the library behind
 * wigner3j() of src/core/physics/wigner_functions.cc:41-71) needs: its Makefile (:86-101) writes it from the
 * probes of src/test_cc_dbl.c:60-100.  These are the answers of gcc 13 on x86-64 for the probes the double
 * precision entry point wig3jj() depends on; the long double / __float128 variants are not built
 * (cfg/wigxjpf_config.h:36-46 defaults them to 0). */
#define WIGXJPF_HAVE_THREAD 1
#define MULTI_WORD_INT_SIZEOF_ITEM 8
#define TRAIL_ARRAY_EMPTY 1
