/*
 * ccp.h -- C ABI of the B200 conjugate-point engine (libccp_b200.so).
 *
 * Drop-in boundary for ONE hot path of JCJaquette/Computing-Conjugate-Points:
 *   propagateManifold::frameDet            (reference propagateManifold.cpp:125-188)
 *     -> computeTotalTrajectory x n        (propagateManifold.cpp:84-121)
 *     -> getTotalTrajectory                (utils.cpp:175-263)
 *     -> topFrame::initialize / countZeros / getLastFrame / getFirstFrame
 *                                          (topFrame.cpp:4-108, 193-351)
 * The reference has no FFI layer (everything is CAPD value types passed between C++
 * member functions), so the boundary is drawn around "n trajectories + topFrame" and
 * expressed as plain-old-data descriptors in / result records out.  One descriptor is
 * exactly the data propagateManifold::construct_InitCondU / construct_A_lin
 * (propagateManifold.cpp:6-81) assemble for one parameter value; one result record is
 * what frameDet consumes afterwards (countZeros output, last/first frame for the L_+ test).
 *
 * No C++ or torch types cross this boundary.  All functions return 0 on success or a
 * negative CCP_E_* code; numerical failures of a single front are reported per front in
 * ccp_front_result.status and never abort the batch (mirrors the per-parameter try/catch
 * of SampleParameters, heteroclinic.cpp:456-467).
 */
#ifndef CCP_H
#define CCP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCP_MAX_N      4      /* coupled PDE components n (reference: 2 or 3; 4 = chain extension) */
#define CCP_MAX_DIM    8      /* 2n: phase-space dimension of the front ODE                          */
#define CCP_MAX_ORDER  20     /* Taylor order p (reference: order = 20, heteroclinic.cpp:63)         */
#define CCP_MAX_GRID   15     /* sub-intervals per step (reference: grid = 14, heteroclinic.cpp:66)  */
#define CCP_MAX_CONJ   8      /* recorded conjugate-point time intervals per front                   */

/* closed interval [lo,hi] of doubles (CAPD `interval` restricted to double end points) */
typedef struct { double lo, hi; } ccp_interval;

/* ---- input: one standing front (one parameter value, or one sub-box of its initial set) ---- */
typedef struct {
    int32_t n;            /* number of components; dimension = 2n, coupled system d = 4n                 */
    int32_t order;        /* Taylor order p          (propagateManifold.h:63 `order`)                   */
    int32_t step_log2;    /* fixed step h = 2^-step_log2 (propagateManifold.cpp:92 `step_size`)         */
    int32_t grid;         /* sub-intervals per step  (frameDet argument `grid`); 0 = flow map only
                             (Gxy / DGxy: no frame tables, grid loop, determinants; flow_P and last_frame only) */
    double  L_minus;      /* integration time L_- as an exact double (heteroclinic.cpp:324)             */
    double  reserved;     /* keeps sizeof(ccp_front_desc) a multiple of 16 (coalesced 16-byte staging)  */
    ccp_interval b[CCP_MAX_N];        /* b_1..b_n  (n=2: a,b)  ODE_functions.cpp:257-273                */
    ccp_interval c[CCP_MAX_N - 1];    /* c_12, c_23, ... chain couplings (n=2: c)                       */
    double  A_u[CCP_MAX_DIM * CCP_MAX_DIM];   /* row-major, leading dim CCP_MAX_DIM: eigenbasis at (0,..,0),
                                                 columns 0..n-1 unstable, n..2n-1 stable (point matrix;
                                                 heteroclinic.cpp:116-117, construct_A_lin)               */
    ccp_interval p_i[CCP_MAX_DIM];    /* phi(-L_-) in global coords (localManifold::getPointNbd [0])    */
    ccp_interval Uxy[CCP_MAX_DIM];    /* its error box in local eigen-coords (getPointNbd [1])          */
    ccp_interval Eu_err[CCP_MAX_N * CCP_MAX_N]; /* row-major ld CCP_MAX_N: Eu_m_Error_Final[j][k] =
                                                 error of eigen-direction k along stable coordinate j
                                                 (localManifold.cpp:673-681)                             */
} ccp_front_desc;

/* per-front status */
enum {
    CCP_OK              = 0,
    CCP_ST_BAD_DESC     = 1,   /* n/order/grid out of range, L_minus <= 0, non-finite input            */
    CCP_ST_ENCLOSURE    = 2,   /* a-priori (rough) enclosure of a step could not be verified            */
    CCP_ST_NONFINITE    = 3    /* NaN/inf appeared in an enclosure                                     */
};

/* ---- output: what frameDet / topFrame hand back for one front ---- */
typedef struct {
    int32_t status;           /* CCP_OK or CCP_ST_*                                                     */
    int32_t zero_count;       /* certified sign changes of det(top frame)  (countZeros output[0])      */
    int32_t failure_count;    /* sub-intervals that could not be certified (countZeros output[1])      */
    int32_t series_length;    /* steps * grid                                                          */
    int32_t steps;            /* integrator steps taken (last one may be partial, utils.cpp:258)       */
    int32_t n_conj;           /* min(zero_count, CCP_MAX_CONJ) entries valid below                      */
    int32_t first_failure;    /* sub-interval index of first failure, -1 if none                       */
    int32_t reserved;
    ccp_interval conj_time[CCP_MAX_CONJ];   /* time label (prevTime+sub) of each certified crossing     */
    int32_t      conj_index[CCP_MAX_CONJ];  /* its sub-interval index                                   */
    /* topFrame::getLastFrame (topFrame.cpp:265-299): rows 0..2n-1, row-major ld = CCP_MAX_N+2.
       cols 0..n-1: (U,V) of trajectory k at the right end of the last sub-interval;
       col n: phi' over the whole last sub-interval; col n+1: phi at the right end.                    */
    ccp_interval last_frame[CCP_MAX_DIM * (CCP_MAX_N + 2)];
    /* data for topFrame::getFirstFrame (topFrame.cpp:301-331), row-major ld = CCP_MAX_N+1.
       cols 0..n-1: (U,V) of trajectory k at the left end of the first sub-interval;
       col n: phi' over the whole first sub-interval.                                                  */
    ccp_interval first_frame[CCP_MAX_DIM * (CCP_MAX_N + 1)];
    ccp_interval det_first;   /* det range enclosure on the first sub-interval                          */
    ccp_interval det_last;    /* det range enclosure on the last sub-interval                           */
    /* C^1 time-L_minus map of the front ODE over the initial set, times the linearisation:
       flow_P = [D Phi_T(set)] * A_u, rows/cols 0..2n-1, row-major ld = CCP_MAX_DIM.  This is `XY_deriv*A_i` of
       boundaryValueProblem::DGxy (boundaryValueProblem.cpp:61-115, C1Rect2Set(p_i,A_i,Uxy)); the C^0 image
       Phi_T(set) of boundaryValueProblem::Gxy (:9-59, C0Rect2Set(p_i,A_i,Uxy)) is last_frame column n+1.
       With grid = 0 (flow mode) last_frame column n holds the image of the CENTRE of the set, Phi_T(p_i): together with flow_P it
       encloses the image of every sub-box U' of Uxy (0 in U') by the mean-value form  Phi_T(p_i) + flow_P U'  -- the point
       image and the box derivative of a Krawczyk step then come from ONE integration (host/ccp_bvp.hpp).
       Does not depend on Eu_err.  The backward maps (STABLE = true, field f_minus) follow from the time-reversal
       symmetry Phi_{-T} = S Phi_T S, S = diag(I,-I) (host/ccp_bvp.hpp).                                        */
    ccp_interval flow_P[CCP_MAX_DIM * CCP_MAX_DIM];
} ccp_front_result;

/* optional per-sub-interval series (topFrame::time_series / det_series, makePlot) */
typedef struct {
    ccp_interval time;    /* prevTime + sub                 (utils.cpp:242)  */
    ccp_interval det_L;   /* det of left-end frame          det_series[i][0] */
    ccp_interval det_R;   /* det of right-end frame         det_series[i][1] */
    ccp_interval det_V;   /* det of range frame             det_series[i][2] (what plot_det.txt holds) */
    ccp_interval det_D;   /* d/dt det, Jacobi formula       calculateDerivative (topFrame.cpp:110-122) */
} ccp_series_rec;

/* error codes of the API itself */
enum {
    CCP_E_CUDA      = -100,   /* CUDA runtime error (see ccp_last_error)            */
    CCP_E_NO_DEVICE = -101,   /* no CUDA device: there is NO CPU fallback           */
    CCP_E_ARG       = -102,
    CCP_E_NOT_INIT  = -103
};

/* ---------------- engine life cycle ---------------- */
int  ccp_device_count(void);
int  ccp_init(int device);            /* binds this process to one GPU, creates stream + scratch.  A second call with the
                                         same device is a no-op; with ANOTHER device it fails with CCP_E_ARG (call
                                         ccp_shutdown first)                                                          */
int  ccp_init_multi(int n_devices);   /* binds devices 0..n_devices-1 (<= 0: every visible GPU): one stream and staging per
                                         device, all driven from the calling host thread.  The reference's sweep is the C++
                                         loop heteroclinic.cpp:449-471; with this call its replacement uses the whole box
                                         (ccp_frame_count_batch_multi).  Single-device entry points run on device 0.        */
int  ccp_bound_devices(void);         /* number of GPUs bound by ccp_init / ccp_init_multi (0 before)                       */
void ccp_shutdown(void);
const char* ccp_last_error(void);
const char* ccp_version(void);

/* ---------------- the hot path ---------------- */
/* Replaces the call `E_u.frameDet(L_minus,grid,endPoint_LPlus-p_s)` (heteroclinic.cpp:334) for a
   batch of fronts: host buffers in, host buffers out (H2D + kernel + D2H inside). */
int ccp_frame_count_batch(const ccp_front_desc* fronts, size_t n_fronts, ccp_front_result* results);

/* The same batch sharded over every bound GPU (ccp_init_multi): front f runs on device f mod n_devices, no exchange between
   devices, the records land in `results` in front order.  With one bound device it is ccp_frame_count_batch.              */
int ccp_frame_count_batch_multi(const ccp_front_desc* fronts, size_t n_fronts, ccp_front_result* results);

/* Same, device-resident: `d_fronts`/`d_results` are device pointers, `stream` a cudaStream_t (or NULL
   for the engine stream); uniform n per batch.  ccp_frame_count_batch_device reads n from the first descriptor (a 4-byte
   device-to-host copy and one stream synchronisation before the launch); ccp_frame_count_batch_device_n takes n from the
   caller and only enqueues.  */
int ccp_frame_count_batch_device(const void* d_fronts, size_t n_fronts, void* d_results, void* stream);
int ccp_frame_count_batch_device_n(const void* d_fronts, size_t n_fronts, void* d_results, void* stream, int n);

/* ---------------- BASELINE config 4: the initial box subdivided, fused propagate + count, hull / tally on the device ----------------
   The box of every front is cut into parts^n sub-boxes with the reference's rule part() (utils.cpp:3-10; index rule of getSubdivision,
   utils.cpp:309-321), each sub-box is propagated and counted independently (same kernel as ccp_frame_count_batch), and the sub-box
   results of one parameter are reduced on the device: the counts are tallied (the parameter is certified iff every sub-box is certified
   and all agree) and every enclosure of the result record is replaced by the interval hull over the sub-boxes.                        */
enum {
    CCP_SUBDIV_UXY    = 0,   /* the x-part of Uxy: local unstable coordinates of the neighbourhood of the connecting orbit (XY_nbd) */
    CCP_SUBDIV_EU_ERR = 1    /* the rows of Eu_m_Error_Final (every column cut the same way)                                         */
};
typedef struct {
    int32_t status;        /* worst status over the sub-boxes (CCP_OK = all ran)                                  */
    int32_t sub_boxes;     /* parts^n                                                                            */
    int32_t certified;     /* sub-boxes with status CCP_OK and failure_count == 0                                */
    int32_t zero_count;    /* the common count if every sub-box is certified and all agree, else -1             */
    int32_t zero_min, zero_max;   /* over the certified sub-boxes (-1 if none)                                  */
    int32_t failures;      /* sum of failure_count                                                               */
    int32_t steps;
    /* interval hulls over the sub-boxes, same layout as ccp_front_result from last_frame on */
    ccp_interval last_frame[CCP_MAX_DIM * (CCP_MAX_N + 2)];
    ccp_interval first_frame[CCP_MAX_DIM * (CCP_MAX_N + 1)];
    ccp_interval det_first, det_last;
    ccp_interval flow_P[CCP_MAX_DIM * CCP_MAX_DIM];
} ccp_subdiv_result;
/* `parts` pieces per coordinate (1..64), `which` = CCP_SUBDIV_*; uniform n over the batch.  Host buffers in / out; the sub-box
   descriptors and records never leave the device.  Fronts are processed in chunks that fit the device buffers (<= 2^17 sub-boxes). */
int ccp_frame_count_subdivided(const ccp_front_desc* fronts, size_t n_fronts, int parts, int which, ccp_subdiv_result* results);

/* One front with the full per-sub-interval series (topFrame::makePlot equivalent).  `cap` records
   available in `series`; *len receives steps*grid.  */
int ccp_frame_count_series(const ccp_front_desc* front, ccp_front_result* result,
                           ccp_series_rec* series, size_t cap, size_t* len);

/* ---------------- next row: localManifold::boundDFU ----------------
   Interval hull, over the N^n sub-boxes of the neighbourhood U, of the derivative of the local vector field
   Ainv * Df(p + A w) * A  (localManifold.cpp:347-409, localVField.h:31-35; part(): utils.cpp:3-10).  One descriptor per call of
   boundDFU; the reference calls it >= 4 times per front with N = 15 (3,375 sub-boxes for n = 3).                              */
#define CCP_DFU_MAX_SUBDIV 32
typedef struct {
    int32_t n;            /* components; the field is 2n-dimensional                                              */
    int32_t subdivision;  /* N = subdivisionNUM (localManifold.h:28; heteroclinic.cpp:64 manifold_subdivision = 15) */
    int32_t stable;       /* 1: subdivide coordinates n..2n-1 (stable manifold), 0: coordinates 0..n-1            */
    int32_t reserved;
    ccp_interval b[CCP_MAX_N];
    ccp_interval c[CCP_MAX_N];                  /* c[0..n-2]                                                        */
    double       A[CCP_MAX_DIM * CCP_MAX_DIM];  /* (*pF).A, point matrix, row-major ld CCP_MAX_DIM                  */
    ccp_interval Ainv[CCP_MAX_DIM * CCP_MAX_DIM]; /* (*pF).Ainv = gaussInverseMatrix(A) (host, stays CAPD)          */
    ccp_interval p[CCP_MAX_DIM];                /* (*pF).p, the equilibrium                                         */
    ccp_interval U[CCP_MAX_DIM];                /* the neighbourhood in local coordinates (constructU)              */
} ccp_dfu_desc;
typedef struct {
    int32_t status;       /* CCP_OK or CCP_ST_BAD_DESC                                                             */
    int32_t parts;        /* N^n                                                                                   */
    ccp_interval DF[CCP_MAX_DIM * CCP_MAX_DIM]; /* the hull, row-major ld CCP_MAX_DIM                              */
} ccp_dfu_result;
int ccp_bound_dfu_batch(const ccp_dfu_desc* descs, size_t n_descs, ccp_dfu_result* results);

/* kernel launch accounting + last kernel time for bench.py */
uint64_t ccp_kernel_launches(void);
float    ccp_last_kernel_ms(void);

/* Unit-test probe: applies primitive `op` of csrc/ccp_interval.cuh element-wise on the device
   (8 input doubles -> 4 output doubles per element); tests compare with the oracle bit for bit. */
int ccp_debug_primitives(int op, const double* in, double* out, size_t n);

/* FP64 DFMA-chain peak of the bound device in TFLOP/s (roofline denominator) */
int ccp_measure_fp64_peak(int iters, double* tflops);

/* FP64 work of one front under the implemented algorithm: flop (1 DFMA = 2) */
double ccp_front_flops(int n, int order, int grid, int steps);

#ifdef __cplusplus
}
#endif
#endif /* CCP_H */
