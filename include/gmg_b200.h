/* gmg_b200.h -- C ABI of the B200-native multigrid-preconditioned CG pressure solve.
 *
 * This is the drop-in boundary for ONE path of rgoldade/GeometricMultigridPressureSolver: the
 * McAdams-2010 MGPCG pressure solve (namespace HDK::GeometricMultigridOperators, class
 * HDK::GeometricMultigridPoissonSolver, HDK::solveGeometricConjugateGradient).  The reference has
 * no FFI of its own for this path -- it is plain in-process C++ called from the Houdini node
 * (HDK_GeometricFreeSurfacePressureSolver.cpp:344-484) -- so each entry point below names the
 * reference function it replaces; include/gmg_b200_hdk.hpp is the C++ facade with the reference's
 * own names/signatures over these calls, and INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Every function returns a gmg_status (0 = ok); there is NO CPU
 *     fallback: without a usable CUDA device every call fails with GMG_ERR_CUDA.
 *   - Host grids are dense, x-fastest: idx = x + res[0]*(y + res[1]*z), in the reference's
 *     EXPANDED coordinates (the power-of-two padded grid of
 *     HDK_GeometricMultigridOperators.h:1340-1360).  Labels are int32 with the enum of
 *     HDK_GeometricMultigridOperators.h:11.  Values are fp64 (the reference is fp64 end to end:
 *     HDK_GeometricMultigridPoissonSolver.h:14-15).  The face grid of axis a has res[a]+1 entries
 *     along a; face (i,j,k) of axis 0 lies between cells (i-1,j,k) and (i,j,k).
 *   - On the device the library stores only the cropped box of non-EXTERIOR cells (+halo) per level;
 *     expanded coordinates stay virtual.
 *   - A solver object is not re-entrant (like the reference's, whose scratch grids are members).
 */
#ifndef GMG_B200_H
#define GMG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum gmg_status
{
    GMG_OK = 0,
    GMG_ERR_CUDA = 1,        /* no device / CUDA runtime error (message via gmg_last_error) */
    GMG_ERR_INVALID = 2,     /* bad argument (null pointer, odd resolution, level out of range ...) */
    GMG_ERR_NO_ACTIVE = 3,   /* no INTERIOR/BOUNDARY cell at level 0 */
    GMG_ERR_COARSE_SIZE = 4, /* coarsest level too large for the dense direct solve */
    GMG_ERR_NOT_SPD = 5,     /* coarse matrix factorisation failed (e.g. pure-Neumann domain) */
    GMG_ERR_COMM = 6         /* multi-GPU exchange failure */
} gmg_status;

/* HDK_GeometricMultigridOperators.h:11 */
enum { GMG_INTERIOR_CELL = 0, GMG_EXTERIOR_CELL = 1, GMG_DIRICHLET_CELL = 2, GMG_BOUNDARY_CELL = 3 };

typedef struct gmg_ctx gmg_ctx;       /* one CUDA device + stream (+ the z-slab communicator when sharded) */
typedef struct gmg_solver gmg_solver; /* GeometricMultigridPoissonSolver */
typedef struct gmg_grid gmg_grid;     /* device-resident fp64 vector grid of one solver level */

const char *gmg_last_error(void);
int gmg_version(void);

/* ---- context ------------------------------------------------------------------------------- */
/* device: CUDA ordinal.  stream: a cudaStream_t to run on (e.g. torch's current stream), or NULL to
 * let the library create its own non-blocking stream. */
int gmg_ctx_create(int device, void *stream, gmg_ctx **out);
int gmg_ctx_destroy(gmg_ctx *ctx);
int gmg_ctx_synchronize(gmg_ctx *ctx);
/* Shard the fine levels into z-slabs over `world` ranks (one process per GPU).  `nccl_unique_id` is the
 * 128-byte ncclUniqueId made on rank 0 and handed to every rank by the caller (torch.distributed /
 * MPI / files -- the library does not care).  Must be called before gmg_solver_create. */
int gmg_ctx_shard(gmg_ctx *ctx, int rank, int world, const void *nccl_unique_id);
int gmg_nccl_unique_id(void *out128);
int gmg_ctx_rank(gmg_ctx *ctx, int *rank, int *world);
/* Sharded contexts (the reference has no counterpart: it is one shared-memory process).  Levels
 * [0, S) are cut into z-slabs with nesting cut planes, one per rank, stored with a deep halo that is recomputed
 * redundantly (two plane exchanges per level per V-cycle instead of one per sweep); levels [S, L) are replicated.
 * Every rank passes the SAME full host grids to gmg_solver_create / gmg_grid_upload / gmg_vcycle / gmg_pcg and keeps
 * its slab; host outputs (gmg_grid_download, x of gmg_vcycle / gmg_pcg) carry the rank's OWNED z-planes only.
 * Scalars (dot products, norms, iteration counts, residual histories) are identical on every rank.
 *
 * gmg_shard_plan is the pure host function behind the partition (no device needed): levelPlanes[l] = z extent of level
 * l's storage box, levelShiftZ[l]: coarse plane = (fine plane >> 1) + shift, levelCells[l] = cells of the box.  It returns
 * the number of sharded levels S (<= maxShardLevels, boxes of >= minCells cells, every rank owning at least the halo
 * depth) and cuts[l*(world+1) + k] = first storage plane of rank k at level l (even; cuts nest between levels). */
int gmg_shard_plan(const int64_t *levelPlanes, const int64_t *levelShiftZ, const int64_t *levelCells, int levels, int world,
		   int maxShardLevels, int64_t minCells, int *shardLevels, int64_t *cuts);

/* ---- domain builders (integer work, bit-exact) --------------------------------------------------- */
/* HDK_GeometricMultigridOperators.h:1340-1360: level count, padding, power-of-two expanded resolution. */
int gmg_expand_dims(const int64_t baseRes[3], int64_t expRes[3], int64_t offset[3], int *mgLevels);
/* buildExpandedCellLabels, HDK_GeometricMultigridOperators.h:1328-1456.  out: host int32[expRes]. */
int gmg_expand_labels(gmg_ctx *ctx, const int32_t *base, const int64_t baseRes[3], int32_t *out,
		      const int64_t expRes[3], const int64_t offset[3]);
/* buildExpandedBoundaryWeights, HDK_GeometricMultigridOperators.h:1458-1572.  out: host double[expRes + 1 along axis]. */
int gmg_expand_weights(gmg_ctx *ctx, const double *baseW, const int64_t baseRes[3], double *out,
		       const int64_t expRes[3], const int64_t offset[3], int axis);
/* setBoundaryCellLabels, HDK_GeometricMultigridOperators.h:1574-1644.  labels rewritten in place (host).
 * boxLo/boxHi (nullable): expanded-coordinate bounds [lo,hi) outside which every label is EXTERIOR
 * (offset and offset+baseRes from gmg_expand_dims) -- spares a host scan. */
int gmg_set_boundary_labels(gmg_ctx *ctx, int32_t *labels, const int64_t res[3], const double *w0,
			    const double *w1, const double *w2, const int64_t boxLo[3], const int64_t boxHi[3]);
/* buildCoarseCellLabels, HDK_GeometricMultigridOperators.cpp:23-163.  coarse: host int32[res/2]. */
int gmg_coarsen_labels(gmg_ctx *ctx, const int32_t *fine, const int64_t fineRes[3], int32_t *coarse);
/* buildBoundaryCells, HDK_GeometricMultigridOperators.cpp:165-469.  Two-call: xyz == NULL returns the count.
 * Order = the reference's sort key (16^3-tile linear index, z, y, x). */
int gmg_boundary_cells(gmg_ctx *ctx, const int32_t *labels, const int64_t res[3], int width, int64_t *xyz,
		       int64_t *count);

/* ---- solver ------------------------------------------------------------------------------------ */
typedef struct gmg_solver_options
{
    int use_gauss_seidel;     /* 0: damped-Jacobi interior smoother (north_star); 1: tiled Gauss-Seidel, the reference's production
				 default (GFS.cpp:463-466) -- single-GPU contexts only */
    int print_stats;          /* like doPrintStats: per-stage CUDA-event timings to stdout */
    int boundary_width;       /* myBoundarySmootherWidth, default 3 (MG.cpp:141) */
    int boundary_iterations;  /* myBoundarySmootherIterations, default 3 (MG.cpp:142) */
    double coarse_matrix_scale; /* 1 = intended algorithm.  J reproduces the reference run with J UT_ThreadedAlgorithm jobs,
				   whose unsplit assembly loop (MG.cpp:334-389) sums every triplet J times. */
    int64_t box_lo[3], box_hi[3]; /* optional non-EXTERIOR bounds hint (all zero = scan the labels) */
    int operators_only;       /* 1: build labels/bands/records only, no coarse factor -- a handle for the stateless operator
				 functions of the facade (gmg_vcycle / gmg_pcg with the V-cycle then fail with GMG_ERR_INVALID) */
    int mixed_precision;      /* 1: gmg_pcg applies the multigrid preconditioner in fp32 (every V-cycle grid and smoother sweep of the
				 levels that run as kernels; the fused coarse cycle stays fp64) inside the fp64 CG -- the reference's own
				 TODO (README.md:34-35).  NOT the reference arithmetic: the residual history drifts (bench.py reports by how
				 much) although the CG itself, its operator and its convergence test stay fp64.  Single-GPU contexts,
				 damped-Jacobi smoother; gmg_vcycle and the operator entry points are unaffected.  Default 0. */
} gmg_solver_options;
void gmg_solver_default_options(gmg_solver_options *opt);

/* GeometricMultigridPoissonSolver::GeometricMultigridPoissonSolver, HDK_GeometricMultigridPoissonSolver.cpp:135-418.
 * Copies what it needs to the device (the labels of the non-EXTERIOR box; of the weights only the six face weights of every
 * level-0 BOUNDARY cell, gathered on the host: nothing else reads a face weight, Ops.h:208-255) -- the host grids may be freed
 * when the call returns, as with the reference's deep copies (MG.cpp:164-180) --, builds per-level labels, boundary bands and
 * the coarse factor.
 * w0 = w1 = w2 = NULL is the reference's `boundaryWeights == nullptr` operator form (weight 1, Ops.h:237-248). */
int gmg_solver_create(gmg_ctx *ctx, const int32_t *labels, const int64_t res[3], const double *w0, const double *w1,
		      const double *w2, int mgLevels, const gmg_solver_options *opt, gmg_solver **out);
/* The same constructor on ONE-BYTE labels (the enum of HDK_GeometricMultigridOperators.h:11 stored as uint8_t): a quarter of
 * the host memory and PCIe traffic of the int32 form; everything else identical. */
int gmg_solver_create_u8(gmg_ctx *ctx, const uint8_t *labels, const int64_t res[3], const double *w0, const double *w1,
			 const double *w2, int mgLevels, const gmg_solver_options *opt, gmg_solver **out);
int gmg_solver_destroy(gmg_solver *s);
/* getMGLevels(), HDK_GeometricMultigridPoissonSolver.h:31 (after the level cap of MG.cpp:243-248) */
int gmg_solver_levels(gmg_solver *s, int *levels);
int gmg_solver_level_res(gmg_solver *s, int level, int64_t res[3]);
/* per-level labels / boundary-band lists in expanded coordinates, for bit-exact checks */
int gmg_solver_get_labels(gmg_solver *s, int level, int32_t *out);
int gmg_solver_get_boundary_cells(gmg_solver *s, int level, int64_t *xyz, int64_t *count);
int gmg_solver_active_cells(gmg_solver *s, int level, int64_t *count); /* over the whole level, also when sharded */
/* sharded: is this level a z-slab; expanded z range [lo, hi) of the rank's owned planes; active cells it stores */
int gmg_solver_shard_info(gmg_solver *s, int level, int *sharded, int64_t *ownLoZ, int64_t *ownHiZ, int64_t *localActive);
int gmg_solver_coarse_unknowns(gmg_solver *s, int64_t *count);
int gmg_solver_setup_ms(gmg_solver *s, double *ms);
/* cells one host<->device transfer of a level-0 vector grid (rhs, initial guess, pressure) moves, and the number of 3D copies it
   is split into: only the bounding rectangles of the active cells of each z-plane travel (everything else is 0 on both sides,
   the reference's "vector grids are zero off the active cells" invariant, HDK_GeometricMultigridOperators.h:756) */
int gmg_solver_transfer_cells(gmg_solver *s, int64_t *cells, int64_t *copies);
/* the pure host function behind that plan (no device needed): extents[z] = (x0, x1, y0, y1), half-open, of the active cells of
   plane z (x1 <= x0: empty plane); groups[k] = (z0, z1, x0, x1, y0, y1), at most `planes` of them; cells = cells the groups hold */
int gmg_transfer_plan(const int32_t *extents, int planes, int32_t *groups, int *groupCount, int64_t *cells);
/* the pure host function behind the constructor's sparse face weights (no device needed; the weight grids of
   buildExpandedBoundaryWeights, HDK_GeometricMultigridOperators.h:1458-1572, are read where they lie): out[n * count + k] = weight
   of face n (-x,+x,-y,+y,-z,+z) of the cell with storage index idx[k] in a box of row pitch `pitch`, plane size `plane` and
   expanded origin org; 0 outside bounds = (lo[3], hi[3]) (null: the whole grids) */
int gmg_gather_face_weights(const int32_t *idx, int64_t count, int pitch, int64_t plane, const int32_t org[3], const double *w0, const double *w1,
			    const double *w2, const int64_t res[3], const int64_t *bounds, double *out);

/* applyVCycle, HDK_GeometricMultigridPoissonSolver.cpp:420-881: host x (in/out), host b. */
int gmg_vcycle(gmg_solver *s, double *x, const double *b, int useInitialGuess);
/* solveGeometricConjugateGradient wired as HDK_GeometricFreeSurfacePressureSolver.cpp:430-483 does
 * (A = applyPoissonMatrix with the fine weights, M^-1 = applyVCycle), HDK_GeometricCGPoissonSolver.h:11-207.
 * x: host in/out (warm start allowed).  iterations: the index CG.h:198 prints, -1 on the two early-outs.
 * relResHistory[k] = sqrt(|r_k|^2/|b|^2), the value CG.h:159 prints; histCount = entries written.
 * preconditioner: 1 = multigrid V-cycle (GFS.cpp:468-483); 2 = the node's other branch, the diagonal preconditioner of
 * HDK_GeometricFreeSurfacePressureSolver.cpp:485-618 (1/6 on INTERIOR cells, 1/(sum of the six face weights) on BOUNDARY cells);
 * 0 = none (plain CG; not a mode of the node).  maxIt <= 0: x untouched, iterations = 0 (CG.h:100).
 * The loop itself runs on the device (a WHILE node around the captured iteration): one host synchronisation per solve. */
int gmg_pcg(gmg_solver *s, double *x, const double *b, double tol, int maxIt, int preconditioner, int *iterations,
	    double *relResHistory, int histCap, int *histCount);
/* The same solve with the initial guess DECLARED zero by the caller -- the node's solutionGrid.constant(0) without a warm start
 * (HDK_GeometricFreeSurfacePressureSolver.cpp:392-398; a constant UT_VoxelArray says so in O(tiles)).  x is output only: it is
 * neither read nor uploaded, the device grid is cleared instead; same arithmetic (r = b - A 0), same results as gmg_pcg on zeros. */
int gmg_pcg_from_zero(gmg_solver *s, double *x, const double *b, double tol, int maxIt, int preconditioner, int *iterations,
		      double *relResHistory, int histCap, int *histCount);

/* ---- device-resident grids and single operators (what the facade's operator functions call) ----------- */
int gmg_grid_create(gmg_solver *s, int level, gmg_grid **out); /* zero-filled */
int gmg_grid_destroy(gmg_grid *g);
int gmg_grid_upload(gmg_grid *g, const double *host);   /* host: dense expanded grid of that level */
int gmg_grid_download(gmg_grid *g, double *host);        /* cells outside the stored box are written as 0 */
int gmg_grid_zero(gmg_grid *g);
int gmg_grid_copy(gmg_grid *dst, const gmg_grid *src);

/* jacobiPoissonSmoother, Ops.h:262-367 (x in place; level 0 uses the fine weights) */
int gmg_jacobi(gmg_solver *s, gmg_grid *x, const gmg_grid *b);
/* tiledGaussSeidelPoissonSmoother, Ops.h:369-520: one half-pass over the odd or even 16^3 tiles, forwards or backwards
 * (x in place).  Needs a solver created with use_gauss_seidel = 1. */
int gmg_gauss_seidel(gmg_solver *s, gmg_grid *x, const gmg_grid *b, int oddTiles, int forward);
/* boundaryJacobiPoissonSmoother over the solver's own band, Ops.h:524-619, `sweeps` times */
int gmg_boundary_jacobi(gmg_solver *s, gmg_grid *x, const gmg_grid *b, int sweeps);
/* applyPoissonMatrix, Ops.h:621-714 (dst written on active cells only) */
int gmg_apply(gmg_solver *s, gmg_grid *dst, const gmg_grid *src);
/* computePoissonResidual, Ops.h:716-732 */
int gmg_residual(gmg_solver *s, gmg_grid *r, const gmg_grid *x, const gmg_grid *b);
/* downsample, Ops.h:734-835 (coarse = level of fine + 1) */
int gmg_restrict(gmg_solver *s, gmg_grid *coarse, const gmg_grid *fine);
/* upsampleAndAdd, Ops.h:873-972 */
int gmg_prolong_add(gmg_solver *s, gmg_grid *fine, const gmg_grid *coarse);
/* dotProduct Ops.h:1020-1085, squaredL2Norm :1205-1265, infNorm :1267-1326 (max(v,0) as in the reference) */
int gmg_dot(gmg_solver *s, const gmg_grid *a, const gmg_grid *b, double *out);
int gmg_norm2(gmg_solver *s, const gmg_grid *a, double *out);
int gmg_inf_norm(gmg_solver *s, const gmg_grid *a, double *out);
/* addToVector Ops.h:1087-1137: dst += scale*src */
int gmg_axpy(gmg_solver *s, gmg_grid *dst, const gmg_grid *src, double scale);
/* addVectors Ops.h:1139-1195: dst = a + scale*v (dst may alias a or v) */
int gmg_add_scaled(gmg_solver *s, gmg_grid *dst, const gmg_grid *a, const gmg_grid *v, double scale);
/* scaleVector Ops.h:974-1018 */
int gmg_scale(gmg_solver *s, gmg_grid *v, double scale);
/* device-resident V-cycle / PCG on grids of level 0 (no host copies; what bench.py's `value` times) */
int gmg_vcycle_device(gmg_solver *s, gmg_grid *x, const gmg_grid *b, int useInitialGuess);
int gmg_pcg_device(gmg_solver *s, gmg_grid *x, const gmg_grid *b, double tol, int maxIt, int preconditioner,
		   int *iterations, double *relResHistory, int histCap, int *histCount);

/* ---- the steps either side of the solve ------------------------------------------------------------
 * The pointwise builders of HDK_GeometricFreeSurfacePressureSolver.cpp that turn the simulation's fields into this path's
 * inputs and its output back into them, on the BASE grid (res = the simulation's cell resolution; cell fields x-fastest, the
 * face field of axis a has one more entry along a).  SIM_RawField values are fpreal32, hence `float`; material labels are the
 * enum of HDK_Utilities.h:17 { SOLID = 0, LIQUID = 1, AIR = 2 } as int32; a face is valid where validFaces == 1
 * (HDK_Utilities.h:21).  rhs / solution are the EXPANDED fp64 grids of the solve (expRes, offset from gmg_expand_dims). */
/* buildMaterialCellLabels, HDK_Utilities.cpp:87-148 (isCellLiquid :5-45): SOLID where none of a cell's six cut-cell weights is > 0;
 * else LIQUID where the surface value is <= 0, or the solid sample is >= 0 and an open face leads to a cell whose surface value
 * is <= 0; else AIR.  solidSurface: the solid SDF sampled at the surface field's cell centres (solidSurface.getValue(pos),
 * HDK_Utilities.cpp:22-24; the HDK interpolation stays with the caller -- for an aligned collision field it is the field). */
int gmg_build_material_labels(gmg_ctx *ctx, const float *liquidSurface, const float *solidSurface, const float *const cutCell[3],
			      const int64_t res[3], int32_t *material);
/* buildValidFaces, GFS.cpp:717-744 (classifyValidFaces, HDK_Utilities.h:137-189), one axis: 1 where the cut-cell weight is > 0 and one
 * of the face's two (in-range) cells is LIQUID, 0 elsewhere */
int gmg_build_valid_faces(gmg_ctx *ctx, const int32_t *material, const float *cutCell, const int64_t res[3], int axis, float *validFaces);
/* buildMGDomainLabels, GFS.cpp:746-793: LIQUID -> INTERIOR, AIR -> DIRICHLET, else EXTERIOR */
int gmg_build_domain_labels(gmg_ctx *ctx, const int32_t *material, const int64_t res[3], int32_t *labels);
/* buildMGBoundaryWeights, GFS.cpp:796-865, one axis: the cut-cell weight on valid faces, divided by the clamped ghost-fluid
 * theta (HDK_Utilities.h:25-42) on a liquid/air face; 0 elsewhere */
int gmg_build_boundary_weights(gmg_ctx *ctx, const float *cutCell, const float *liquidSurface, const float *validFaces,
			       const int32_t *domainLabels, const int64_t res[3], int axis, double *weights);
/* buildRHS, GFS.cpp:868-943: cut-cell divergence of every LIQUID cell, written into the expanded rhs grid (other cells are left
 * as they are: pass it zero-filled like the reference's rhsGrid.constant(0)).  solidVelocity: NULL, or per axis the solid's
 * velocity already sampled at the face centres (the reference samples a SIM_VectorField there, GFS.cpp:918-921). */
int gmg_build_rhs(gmg_ctx *ctx, const int32_t *material, const float *const velocity[3], const float *const cutCell[3],
		  const float *const solidVelocity[3], const int64_t res[3], const int64_t expRes[3], const int64_t offset[3], double *rhs);
/* applyOldPressure, GFS.cpp:946-997 (the warm start) and applySolutionToPressure, GFS.cpp:1000-1047 */
int gmg_apply_old_pressure(gmg_ctx *ctx, const float *pressure, const int32_t *material, const int64_t res[3], const int64_t expRes[3],
			   const int64_t offset[3], double *solution);
int gmg_apply_solution_to_pressure(gmg_ctx *ctx, float *pressure, const int32_t *material, const double *solution, const int64_t res[3],
				   const int64_t expRes[3], const int64_t offset[3]);
/* applyPressureGradient, GFS.cpp:1050-1131, one axis: velocity -= grad p on valid faces, ghost-fluid scaled at the free surface */
int gmg_apply_pressure_gradient(gmg_ctx *ctx, float *velocity, const float *liquidSurface, const float *pressure, const float *validFaces,
				const int32_t *material, const int64_t res[3], int axis);

/* ---- measurement hooks ----------------------------------------------------------------------------- */
/* kernels launched by this library on this context since the last reset (bench.py's gpu_launches) */
int gmg_launch_count(gmg_ctx *ctx, int64_t *count, int reset);
/* communication micro-benchmark on a sharded solver (collective): `reps` back-to-back operations with no compute between them.
 * kind 0 = halo exchange of `depth` planes at `level`, 1 = gather of the first replicated level, 2 = scalar all-reduce. */
int gmg_comm_benchmark(gmg_solver *s, int kind, int level, int depth, int reps, double *msPerOp);
/* communication operations (halo exchanges, gathers, scalar all-reduces) enqueued since the last gmg_launch_count reset */
int gmg_comm_count(gmg_ctx *ctx, int64_t *count);
/* CUDA-event timing on the context's stream: begin/end bracket, result in ms */
int gmg_timer_begin(gmg_ctx *ctx);
int gmg_timer_end(gmg_ctx *ctx, double *ms);
/* accumulated device time (ms) and launches per kernel class since the last reset; classes are listed by
 * gmg_kernel_class_name(i), i in [0, gmg_kernel_class_count()).  Only collected when enabled (adds events).
 * fineLevelOnly != 0 restricts the sums to launches on level 0, where the HBM roofline is quoted.
 * on = 1: one event pair per launch.  on = 2: the same, except that the back-to-back sweeps of a band sweep group share ONE pair
 * (counted as that many launches): an event-record node costs ~5 us and keeps the launches it separates from overlapping, which
 * overstates a 10 us kernel by half. */
int gmg_profile_enable(gmg_ctx *ctx, int on);
int gmg_kernel_class_count(void);
const char *gmg_kernel_class_name(int i);
int gmg_profile_get(gmg_ctx *ctx, int klass, int fineLevelOnly, double *ms, int64_t *launches, double *algorithmicBytes);
int gmg_profile_reset(gmg_ctx *ctx);
/* the same device times split by multigrid level (level < 16) */
int gmg_profile_get_level(gmg_ctx *ctx, int klass, int level, double *ms, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif
