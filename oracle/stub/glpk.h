/* TEST INFRASTRUCTURE ONLY.
 * Stand-in for <glpk.h> so the unmodified reference main.cpp compiles here without GLPK.
 * GLPK is only reached by `kmercamel maskopt -t min-run` (reference src/masks.h:181-237), which is
 * not on the `compute` path; every entry point aborts if it is ever called. */
#pragma once
#include <cstdlib>
#include <cstdio>
struct glp_prob;
enum { GLP_MIN = 1, GLP_LO = 2, GLP_IV = 2, GLP_OFF = 0 };
static inline void kc_glpk_absent() { std::fprintf(stderr, "GLPK is not available in the oracle build\n"); std::abort(); }
static inline glp_prob *glp_create_prob() { kc_glpk_absent(); return nullptr; }
static inline void glp_set_obj_dir(glp_prob *, int) { kc_glpk_absent(); }
static inline int glp_add_rows(glp_prob *, int) { kc_glpk_absent(); return 0; }
static inline int glp_add_cols(glp_prob *, int) { kc_glpk_absent(); return 0; }
static inline void glp_set_row_bnds(glp_prob *, int, int, double, double) { kc_glpk_absent(); }
static inline void glp_set_col_bnds(glp_prob *, int, int, double, double) { kc_glpk_absent(); }
static inline void glp_set_col_kind(glp_prob *, int, int) { kc_glpk_absent(); }
static inline void glp_set_obj_coef(glp_prob *, int, double) { kc_glpk_absent(); }
static inline int glp_term_out(int) { kc_glpk_absent(); return 0; }
static inline void glp_load_matrix(glp_prob *, int, const int *, const int *, const double *) { kc_glpk_absent(); }
static inline int glp_simplex(glp_prob *, const void *) { kc_glpk_absent(); return 0; }
static inline double glp_get_col_prim(glp_prob *, int) { kc_glpk_absent(); return 0; }
static inline void glp_delete_prob(glp_prob *) { kc_glpk_absent(); }
