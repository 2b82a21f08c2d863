/*
 * ccp_setup.h -- C ABI of libccp_setup.so: HOST-ONLY generator of front descriptors (no CUDA, no GPU).
 *
 * Synthetic sweeps need descriptors (p_i, A_u, L_-, boxes) that the reference obtains from stage (i) of
 * computeFrontAndConjugatePoints (heteroclinic.cpp:96-191, 324) with CAPD.  This library re-states that stage in plain
 * doubles (non-rigorous float shooting) and the reference's sweep generator Construct_ParameterList
 * (heteroclinic.cpp:386-427).  It is NOT part of the timed path and lives in its own shared object so that CPU-only
 * consumers (bench.py --impl reference, the oracle tests) never map the CUDA library libccp_b200.so.
 */
#ifndef CCP_SETUP_H
#define CCP_SETUP_H

#include "ccp.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    double xy_nbd;        /* +- radius of the validated BVP neighbourhood in x (local coords); <0 -> 1e-12
                             (the Krawczyk stage of host/ccp_bvp.hpp delivers 1e-13 .. 3e-13)  */
    double cone_L;        /* cone constant L (heteroclinic.cpp:80-87); <=0 -> reference default per n   */
    double eig_err;       /* +- magnitude of Eu_m_Error_Final entries; <0 -> 2e-4 (n=2) / 2e-5 (n>=3)   */
    double scale;         /* manifold scale; <=0 -> reference default per n                             */
    double L_minus_percent; /* <=0 -> 0.84 (heteroclinic.cpp:72)                                        */
    double L_minus_override; /* >0 -> use this L_minus instead of sup(2*T*percent)                      */
    int32_t order, step_log2, grid;   /* <=0 -> 20, 5, 14                                               */
    int32_t newton_steps;             /* <=0 -> 20                                                      */
} ccp_setup_opts;

typedef struct {
    double T;             /* shooting time after FindTime (boundaryValueProblem.cpp:118)                */
    double L_minus;
    double newton_residual;
    double eig_unstable[CCP_MAX_N];
    double x_local[CCP_MAX_N];
    double y_local[CCP_MAX_N];
    int32_t converged;
    int32_t reserved;
} ccp_setup_info;

/* params = (b_1..b_n, c_12..c_{n-1,n}) as doubles */
int ccp_setup_front(int n, const double* params, const ccp_setup_opts* opts,
                    ccp_front_desc* out, ccp_setup_info* info);
/* n_fronts parameter vectors, each 2n-1 doubles, `threads` host threads (<=0: all cores) */
int ccp_setup_sweep(int n, const double* params, size_t n_fronts, const ccp_setup_opts* opts,
                    ccp_front_desc* out, ccp_setup_info* info, int threads);
/* the reference's own sweep generator, Construct_ParameterList (heteroclinic.cpp:386-427):
   writes N_c*N_r vectors (b1,b2,b3,c12,c23) */
int ccp_parameter_list(int N_c, int N_r, double r_min, double r_max, const double* b3, double* out);

#ifdef __cplusplus
}
#endif
#endif /* CCP_SETUP_H */
