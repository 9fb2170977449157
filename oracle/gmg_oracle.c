/* gmg_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement ("port") of the reference's hot path: the McAdams-2010
 * multigrid-preconditioned CG pressure solve of rgoldade/GeometricMultigridPressureSolver.
 * It exists to CHECK the CUDA product; it is never shipped, linked or called by it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Parity status: PINNED.  Every function here is compared, in tests/test_oracle_vs_reference.py,
 * against the reference's own sources compiled unmodified (oracle/_ref/libgmg_ref.so, built by
 * oracle/Makefile from /root/reference/Source + oracle/shim) and against committed fixtures that
 * library generated (tests/golden/, tests/golden/make_golden.py).  The one un-pinned boundary is
 * the coarsest-level direct solve: the reference calls Eigen::SimplicialCholesky (Eigen is
 * un-vendored and version-unpinned: README.md:9, cmake/FindEIGEN3.cmake:20-31) and stores no
 * values for it; any exact fp64 factorisation agrees to round-off.
 *
 * All grids are dense, x-fastest (idx = x + rx*(y + ry*z)), in the reference's expanded
 * coordinates.  File abbreviations: Ops.h/.cpp = Source/HDK_GeometricMultigridOperators.{h,cpp},
 * MG.cpp = Source/HDK_GeometricMultigridPoissonSolver.cpp, CG.h = Source/HDK_GeometricCGPoissonSolver.h.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* Ops.h:11 */
enum { INTERIOR_CELL = 0, EXTERIOR_CELL = 1, DIRICHLET_CELL = 2, BOUNDARY_CELL = 3 };

#define TILE 16 /* UT_VoxelArray tile edge */

typedef int64_t i64;

static inline i64 lin(const i64 res[3], i64 x, i64 y, i64 z) { return x + res[0] * (y + res[1] * z); }
static inline int is_active(int l) { return l == INTERIOR_CELL || l == BOUNDARY_CELL; }
static inline i64 cells_of(const i64 res[3]) { return res[0] * res[1] * res[2]; }

/* clamped label read, like UT_VoxelArray::operator() */
static inline int label_at(const int *labels, const i64 res[3], i64 x, i64 y, i64 z)
{
    x = x < 0 ? 0 : (x >= res[0] ? res[0] - 1 : x);
    y = y < 0 ? 0 : (y >= res[1] ? res[1] - 1 : y);
    z = z < 0 ? 0 : (z >= res[2] ? res[2] - 1 : z);
    return labels[lin(res, x, y, z)];
}

int orc_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ domain builders */

/* Ops.h:1340-1360: level count, padding and power-of-two expanded resolution */
void orc_expand_dims(const i64 res[3], i64 outRes[3], i64 offset[3], int *mgLevels)
{
    double minLog = fmin(log2((double)res[0]), log2((double)res[1]));
    minLog = fmin(minLog, log2((double)res[2]));
    int levels = (int)(ceil(minLog) - log2(2.0));
    int pad = (int)pow(2.0, levels - 1);
    for (int a = 0; a < 3; ++a)
    {
	double logSize = ceil(log2((double)(res[a] + 2 * pad)));
	outRes[a] = (i64)exp2(logSize);
	offset[a] = pad;
    }
    *mgLevels = levels;
}

/* Ops.h:1362-1453: EXTERIOR fill, copy INTERIOR/DIRICHLET at +offset */
void orc_expand_labels(const int *base, const i64 res[3], int *out, const i64 outRes[3], const i64 offset[3])
{
    const i64 n = cells_of(outRes);
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i) out[i] = EXTERIOR_CELL;
#pragma omp parallel for
    for (i64 z = 0; z < res[2]; ++z)
	for (i64 y = 0; y < res[1]; ++y)
	    for (i64 x = 0; x < res[0]; ++x)
	    {
		int l = base[lin(res, x, y, z)];
		if (l == EXTERIOR_CELL) continue;
		out[lin(outRes, x + offset[0], y + offset[1], z + offset[2])] = (l == INTERIOR_CELL) ? INTERIOR_CELL : DIRICHLET_CELL;
	    }
}

/* Ops.h:1458-1572: zero fill, copy weights > 0 to face + offset. Face grids have +1 along `axis`. */
void orc_expand_weights(const double *baseW, const i64 baseRes[3], const i64 expRes[3], const i64 offset[3], int axis, double *outW)
{
    i64 bfr[3] = {baseRes[0], baseRes[1], baseRes[2]}, efr[3] = {expRes[0], expRes[1], expRes[2]};
    ++bfr[axis];
    ++efr[axis];
    const i64 n = cells_of(efr);
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i) outW[i] = 0;
#pragma omp parallel for
    for (i64 z = 0; z < bfr[2]; ++z)
	for (i64 y = 0; y < bfr[1]; ++y)
	    for (i64 x = 0; x < bfr[0]; ++x)
	    {
		double w = baseW[lin(bfr, x, y, z)];
		if (w > 0) outW[lin(efr, x + offset[0], y + offset[1], z + offset[2])] = w;
	    }
}

static inline double face_weight(const double *const w[3], const i64 res[3], i64 x, i64 y, i64 z, int axis, int dir)
{
    /* cellToFaceMap: backward face shares the cell index, forward face is +1 along axis */
    i64 fr[3] = {res[0], res[1], res[2]};
    ++fr[axis];
    i64 f[3] = {x, y, z};
    f[axis] += dir;
    return w[axis][lin(fr, f[0], f[1], f[2])];
}

/* Ops.h:1574-1644: INTERIOR -> BOUNDARY if a 6-neighbour is DIRICHLET/EXTERIOR or a face weight != 1.
 * Reads only test for DIRICHLET/EXTERIOR (never rewritten), so in place == out of place. */
void orc_set_boundary_labels(int *labels, const i64 res[3], const double *w0, const double *w1, const double *w2)
{
    const double *const w[3] = {w0, w1, w2};
    const i64 n = cells_of(res);
    unsigned char *promote = (unsigned char *)calloc((size_t)n, 1);
#pragma omp parallel for
    for (i64 z = 0; z < res[2]; ++z)
	for (i64 y = 0; y < res[1]; ++y)
	    for (i64 x = 0; x < res[0]; ++x)
	    {
		if (labels[lin(res, x, y, z)] != INTERIOR_CELL) continue;
		int isBoundary = 0;
		for (int axis = 0; axis < 3 && !isBoundary; ++axis)
		    for (int dir = 0; dir < 2; ++dir)
		    {
			i64 c[3] = {x, y, z};
			c[axis] += dir ? 1 : -1;
			int l = label_at(labels, res, c[0], c[1], c[2]);
			if (l == DIRICHLET_CELL || l == EXTERIOR_CELL) { isBoundary = 1; break; }
			if (face_weight(w, res, x, y, z, axis, dir) != 1) { isBoundary = 1; break; }
		    }
		promote[lin(res, x, y, z)] = (unsigned char)isBoundary;
	    }
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i)
	if (promote[i]) labels[i] = BOUNDARY_CELL;
    free(promote);
}

/* Ops.cpp:23-163 */
void orc_coarsen_labels(const int *fine, const i64 fineRes[3], int *coarse)
{
    const i64 cres[3] = {fineRes[0] / 2, fineRes[1] / 2, fineRes[2] / 2};
#pragma omp parallel for
    for (i64 z = 0; z < cres[2]; ++z)
	for (i64 y = 0; y < cres[1]; ++y)
	    for (i64 x = 0; x < cres[0]; ++x)
	    {
		int hasDirichlet = 0, hasInterior = 0;
		for (int dz = 0; dz < 2; ++dz)
		    for (int dy = 0; dy < 2; ++dy)
			for (int dx = 0; dx < 2; ++dx)
			{
			    int l = fine[lin(fineRes, 2 * x + dx, 2 * y + dy, 2 * z + dz)];
			    if (l == DIRICHLET_CELL) hasDirichlet = 1;
			    else if (is_active(l)) hasInterior = 1;
			}
		coarse[lin(cres, x, y, z)] = hasDirichlet ? DIRICHLET_CELL : (hasInterior ? INTERIOR_CELL : EXTERIOR_CELL);
	    }
    /* Ops.cpp:107-158: INTERIOR -> BOUNDARY next to EXTERIOR/DIRICHLET (tests never look at INTERIOR/BOUNDARY) */
    const i64 n = cells_of(cres);
    unsigned char *promote = (unsigned char *)calloc((size_t)n, 1);
#pragma omp parallel for
    for (i64 z = 0; z < cres[2]; ++z)
	for (i64 y = 0; y < cres[1]; ++y)
	    for (i64 x = 0; x < cres[0]; ++x)
	    {
		if (coarse[lin(cres, x, y, z)] != INTERIOR_CELL) continue;
		int hasBoundary = 0;
		for (int axis = 0; axis < 3 && !hasBoundary; ++axis)
		    for (int dir = 0; dir < 2; ++dir)
		    {
			i64 c[3] = {x, y, z};
			c[axis] += dir ? 1 : -1;
			int l = label_at(coarse, cres, c[0], c[1], c[2]);
			if (l == EXTERIOR_CELL || l == DIRICHLET_CELL) { hasBoundary = 1; break; }
		    }
		promote[lin(cres, x, y, z)] = (unsigned char)hasBoundary;
	    }
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i)
	if (promote[i]) coarse[i] = BOUNDARY_CELL;
    free(promote);
}

/* Ops.cpp:165-469: layer 0 = BOUNDARY cells; layer k+1 = unvisited INTERIOR 6-neighbours of layer k;
 * result = visited cells ordered by (16^3-tile linear index, z, y, x) (Ops.cpp:441-466).
 * Walking tiles in linear order and voxels z,y,x inside a tile yields exactly that order.
 * xyz may be NULL to query the count. Returns the count. */
i64 orc_boundary_cells(const int *labels, const i64 res[3], int width, i64 *xyz, i64 cap)
{
    const i64 n = cells_of(res);
    unsigned char *visited = (unsigned char *)calloc((size_t)n, 1);
    unsigned char *next = (unsigned char *)calloc((size_t)n, 1);
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i) visited[i] = (labels[i] == BOUNDARY_CELL);
    for (int layer = 0; layer < width - 1; ++layer)
    {
#pragma omp parallel for
	for (i64 z = 0; z < res[2]; ++z)
	    for (i64 y = 0; y < res[1]; ++y)
		for (i64 x = 0; x < res[0]; ++x)
		{
		    const i64 i = lin(res, x, y, z);
		    next[i] = visited[i];
		    if (visited[i] || labels[i] != INTERIOR_CELL) continue;
		    int hit = 0;
		    if (x > 0 && visited[i - 1]) hit = 1;
		    if (x < res[0] - 1 && visited[i + 1]) hit = 1;
		    if (y > 0 && visited[i - res[0]]) hit = 1;
		    if (y < res[1] - 1 && visited[i + res[0]]) hit = 1;
		    if (z > 0 && visited[i - res[0] * res[1]]) hit = 1;
		    if (z < res[2] - 1 && visited[i + res[0] * res[1]]) hit = 1;
		    next[i] = (unsigned char)hit;
		}
	unsigned char *t = visited;
	visited = next;
	next = t;
    }
    const i64 tr[3] = {(res[0] + TILE - 1) / TILE, (res[1] + TILE - 1) / TILE, (res[2] + TILE - 1) / TILE};
    i64 count = 0;
    for (i64 tz = 0; tz < tr[2]; ++tz)
	for (i64 ty = 0; ty < tr[1]; ++ty)
	    for (i64 tx = 0; tx < tr[0]; ++tx)
		for (i64 z = tz * TILE; z < (tz + 1) * TILE && z < res[2]; ++z)
		    for (i64 y = ty * TILE; y < (ty + 1) * TILE && y < res[1]; ++y)
			for (i64 x = tx * TILE; x < (tx + 1) * TILE && x < res[0]; ++x)
			    if (visited[lin(res, x, y, z)])
			    {
				if (xyz && count < cap) { xyz[3 * count] = x; xyz[3 * count + 1] = y; xyz[3 * count + 2] = z; }
				++count;
			    }
    free(visited);
    free(next);
    return count;
}

/* ------------------------------------------------------------------ invariant checkers */

/* Ops.h:1771-1870 */
int orc_unit_test_boundary_cells(const int *labels, const i64 res[3], const double *w0, const double *w1, const double *w2)
{
    const double *const w[3] = {w0, w1, w2};
    int ok = 1;
#pragma omp parallel for
    for (i64 z = 0; z < res[2]; ++z)
	for (i64 y = 0; y < res[1]; ++y)
	    for (i64 x = 0; x < res[0]; ++x)
	    {
		int l = labels[lin(res, x, y, z)];
		if (l == INTERIOR_CELL)
		{
		    for (int axis = 0; axis < 3; ++axis)
			for (int dir = 0; dir < 2; ++dir)
			{
			    i64 c[3] = {x, y, z};
			    c[axis] += dir ? 1 : -1;
			    if (!is_active(label_at(labels, res, c[0], c[1], c[2]))) ok = 0;
			}
		}
		else if (l == BOUNDARY_CELL)
		{
		    int valid = 0;
		    for (int axis = 0; axis < 3; ++axis)
			for (int dir = 0; dir < 2; ++dir)
			{
			    i64 c[3] = {x, y, z};
			    c[axis] += dir ? 1 : -1;
			    int nl = label_at(labels, res, c[0], c[1], c[2]);
			    if (!is_active(nl)) valid = 1;
			    else if (w0 && face_weight(w, res, x, y, z, axis, dir) != 1 && nl == BOUNDARY_CELL) valid = 1;
			}
		    if (!valid) ok = 0;
		}
	    }
    return ok;
}

/* Ops.cpp:602-632 */
int orc_unit_test_exterior_cells(const int *labels, const i64 res[3])
{
    int ok = 1;
    for (i64 z = 0; z < res[2]; ++z)
	for (i64 y = 0; y < res[1]; ++y)
	    for (i64 x = 0; x < res[0]; ++x)
		if (x == 0 || y == 0 || z == 0 || x == res[0] - 1 || y == res[1] - 1 || z == res[2] - 1)
		    if (labels[lin(res, x, y, z)] != EXTERIOR_CELL) ok = 0;
    return ok;
}

/* Ops.cpp:471-600 */
int orc_unit_test_coarsening(const int *coarse, const int *fine, const i64 fineRes[3])
{
    const i64 cres[3] = {fineRes[0] / 2, fineRes[1] / 2, fineRes[2] / 2};
    for (int a = 0; a < 3; ++a)
	if (fineRes[a] % 2 || cres[a] % 2 || 2 * cres[a] != fineRes[a]) return 0;
    int ok = 1;
#pragma omp parallel for
    for (i64 z = 0; z < fineRes[2]; ++z)
	for (i64 y = 0; y < fineRes[1]; ++y)
	    for (i64 x = 0; x < fineRes[0]; ++x)
	    {
		int fl = fine[lin(fineRes, x, y, z)];
		int cl = coarse[lin(cres, x / 2, y / 2, z / 2)];
		if (fl == DIRICHLET_CELL && cl != DIRICHLET_CELL) ok = 0;
		else if (is_active(fl) && cl == EXTERIOR_CELL) ok = 0;
	    }
#pragma omp parallel for
    for (i64 z = 0; z < cres[2]; ++z)
	for (i64 y = 0; y < cres[1]; ++y)
	    for (i64 x = 0; x < cres[0]; ++x)
	    {
		int d = 0, in = 0, ex = 0;
		for (int k = 0; k < 8; ++k)
		{
		    int fl = fine[lin(fineRes, 2 * x + (k & 1), 2 * y + ((k >> 1) & 1), 2 * z + ((k >> 2) & 1))];
		    if (fl == DIRICHLET_CELL) d = 1;
		    else if (is_active(fl)) in = 1;
		    else ex = 1;
		}
		int cl = coarse[lin(cres, x, y, z)];
		if (cl == DIRICHLET_CELL) { if (!d) ok = 0; }
		else if (is_active(cl)) { if (d || !in) ok = 0; }
		else { if (d || in || !ex) ok = 0; }
	    }
    return ok;
}

/* ------------------------------------------------------------------ the 7-point operator */

/* Ops.h:177-260 computeLaplacian: returns laplacian, writes diagonal.  Accumulation order is
 * axis-major, direction-minor, centre term last, as in the reference. */
static inline double laplacian_at(const double *u, const int *labels, const i64 res[3], const double *const w[3],
				  i64 x, i64 y, i64 z, double *diagOut)
{
    const i64 i = lin(res, x, y, z);
    const i64 stride[3] = {1, res[0], res[0] * res[1]};
    double lap = 0, diag = 0;
    if (labels[i] == INTERIOR_CELL)
    {
	for (int axis = 0; axis < 3; ++axis)
	    for (int dir = 0; dir < 2; ++dir)
		lap -= u[i + (dir ? stride[axis] : -stride[axis])];
	diag = 6;
    }
    else
    {
	for (int axis = 0; axis < 3; ++axis)
	    for (int dir = 0; dir < 2; ++dir)
	    {
		const i64 j = i + (dir ? stride[axis] : -stride[axis]);
		const int nl = labels[j];
		if (nl == INTERIOR_CELL) { lap -= u[j]; ++diag; }
		else if (nl == BOUNDARY_CELL)
		{
		    if (w[0]) { double wt = face_weight(w, res, x, y, z, axis, dir); lap -= wt * u[j]; diag += wt; }
		    else { lap -= u[j]; ++diag; }
		}
		else if (nl == DIRICHLET_CELL)
		{
		    if (w[0]) diag += face_weight(w, res, x, y, z, axis, dir);
		    else ++diag;
		}
	    }
    }
    lap += diag * u[i];
    *diagOut = diag;
    return lap;
}

/* Ops.h:262-367: snapshot copy, then x += (2/3)(b - A tmp)/diag on active cells */
void orc_jacobi(double *x, const double *b, const int *labels, const i64 res[3], const double *w0, const double *w1, const double *w2)
{
    const double *const w[3] = {w0, w1, w2};
    const i64 n = cells_of(res);
    double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(tmp, x, sizeof(double) * (size_t)n);
    const double damped = 2. / 3.;
#pragma omp parallel for
    for (i64 z = 0; z < res[2]; ++z)
	for (i64 y = 0; y < res[1]; ++y)
	    for (i64 xx = 0; xx < res[0]; ++xx)
	    {
		const i64 i = lin(res, xx, y, z);
		if (!is_active(labels[i])) continue;
		double diag;
		double lap = laplacian_at(tmp, labels, res, w, xx, y, z, &diag);
		double r = b[i] - lap;
		r /= diag;
		x[i] = x[i] + damped * r;
	    }
    free(tmp);
}

/* Ops.h:369-520: tiles with odd/even (tx+ty+tz); lexicographic (x fastest) forward or reverse inside a tile; undamped */
void orc_gauss_seidel(double *x, const double *b, const int *labels, const i64 res[3], int oddTiles, int forward,
		      const double *w0, const double *w1, const double *w2)
{
    const double *const w[3] = {w0, w1, w2};
    const i64 tr[3] = {(res[0] + TILE - 1) / TILE, (res[1] + TILE - 1) / TILE, (res[2] + TILE - 1) / TILE};
    const i64 nt = tr[0] * tr[1] * tr[2];
#pragma omp parallel for schedule(dynamic, 4)
    for (i64 t = 0; t < nt; ++t)
    {
	const i64 tx = t % tr[0], ty = (t / tr[0]) % tr[1], tz = t / (tr[0] * tr[1]);
	const int isOdd = ((tx + ty + tz) % 2) != 0;
	if ((oddTiles && !isOdd) || (!oddTiles && isOdd)) continue;
	const i64 s[3] = {tx * TILE, ty * TILE, tz * TILE};
	i64 e[3] = {s[0] + TILE, s[1] + TILE, s[2] + TILE};
	for (int a = 0; a < 3; ++a) if (e[a] > res[a]) e[a] = res[a];
	const i64 nx = e[0] - s[0], ny = e[1] - s[1], nz = e[2] - s[2];
	const i64 cnt = nx * ny * nz;
	for (i64 k = 0; k < cnt; ++k)
	{
	    const i64 kk = forward ? k : cnt - 1 - k;
	    const i64 cx = s[0] + kk % nx, cy = s[1] + (kk / nx) % ny, cz = s[2] + kk / (nx * ny);
	    const i64 i = lin(res, cx, cy, cz);
	    if (!is_active(labels[i])) continue;
	    double diag;
	    double lap = laplacian_at(x, labels, res, w, cx, cy, cz, &diag);
	    double r = b[i] - lap;
	    r /= diag;
	    x[i] = x[i] + r;
	}
    }
}

/* Ops.h:524-619: Jacobi over the listed cells only, two phases */
void orc_boundary_jacobi(double *x, const double *b, const int *labels, const i64 res[3], const i64 *cells, i64 count,
			 int sweeps, const double *w0, const double *w1, const double *w2)
{
    const double *const w[3] = {w0, w1, w2};
    const double damped = 2. / 3.;
    double *tmp = (double *)malloc(sizeof(double) * (size_t)(count + 1));
    for (int s = 0; s < sweeps; ++s)
    {
#pragma omp parallel for
	for (i64 k = 0; k < count; ++k)
	{
	    const i64 cx = cells[3 * k], cy = cells[3 * k + 1], cz = cells[3 * k + 2];
	    const i64 i = lin(res, cx, cy, cz);
	    double diag;
	    double lap = laplacian_at(x, labels, res, w, cx, cy, cz, &diag);
	    double r = b[i] - lap;
	    r /= diag;
	    tmp[k] = x[i] + damped * r;
	}
#pragma omp parallel for
	for (i64 k = 0; k < count; ++k)
	    x[lin(res, cells[3 * k], cells[3 * k + 1], cells[3 * k + 2])] = tmp[k];
    }
    free(tmp);
}

/* Ops.h:621-714: dst = A src on active cells, dst untouched elsewhere */
void orc_apply(double *dst, const double *src, const int *labels, const i64 res[3], const double *w0, const double *w1, const double *w2)
{
    const double *const w[3] = {w0, w1, w2};
#pragma omp parallel for
    for (i64 z = 0; z < res[2]; ++z)
	for (i64 y = 0; y < res[1]; ++y)
	    for (i64 x = 0; x < res[0]; ++x)
	    {
		const i64 i = lin(res, x, y, z);
		if (!is_active(labels[i])) continue;
		double diag;
		dst[i] = laplacian_at(src, labels, res, w, x, y, z, &diag);
	    }
}

/* Ops.h:716-732: r = 0; r = A x; r = b + (-1) r   (on active cells) */
void orc_residual(double *r, const double *x, const double *b, const int *labels, const i64 res[3], const double *w0, const double *w1, const double *w2)
{
    const i64 n = cells_of(res);
    memset(r, 0, sizeof(double) * (size_t)n);
    orc_apply(r, x, labels, res, w0, w1, w2);
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i)
	if (is_active(labels[i])) r[i] = b[i] + (-1.0) * r[i];
}

/* Ops.h:734-835: coarse (active) = sum_{z,y,x in 0..3} w[x] w[y] w[z] fine(2c-1+(x,y,z)); dest zeroed first */
void orc_downsample(double *coarse, const double *fine, const int *coarseLabels, const i64 fineRes[3])
{
    static const double rw[4] = {1. / 8., 3. / 8., 3. / 8., 1. / 8.};
    const i64 cres[3] = {fineRes[0] / 2, fineRes[1] / 2, fineRes[2] / 2};
    memset(coarse, 0, sizeof(double) * (size_t)cells_of(cres));
#pragma omp parallel for
    for (i64 z = 0; z < cres[2]; ++z)
	for (i64 y = 0; y < cres[1]; ++y)
	    for (i64 x = 0; x < cres[0]; ++x)
	    {
		const i64 c = lin(cres, x, y, z);
		if (!is_active(coarseLabels[c])) continue;
		double v = 0;
		for (int dz = 0; dz < 4; ++dz)
		    for (int dy = 0; dy < 4; ++dy)
			for (int dx = 0; dx < 4; ++dx)
			    v += rw[dx] * rw[dy] * rw[dz] * fine[lin(fineRes, 2 * x - 1 + dx, 2 * y - 1 + dy, 2 * z - 1 + dz)];
		coarse[c] = v;
	    }
}

/* Ops.h:841-871 */
static inline double lerp1(double v0, double v1, double f) { return (1. - f) * v0 + f * v1; }

/* Ops.h:873-972: fine (active) += 4 * trilerp of the 8 coarse cells around p = .5(c+.5)-.5 */
void orc_upsample_add(double *fine, const double *coarse, const int *fineLabels, const i64 fineRes[3])
{
    const i64 cres[3] = {fineRes[0] / 2, fineRes[1] / 2, fineRes[2] / 2};
#pragma omp parallel for
    for (i64 z = 0; z < fineRes[2]; ++z)
	for (i64 y = 0; y < fineRes[1]; ++y)
	    for (i64 x = 0; x < fineRes[0]; ++x)
	    {
		const i64 i = lin(fineRes, x, y, z);
		if (!is_active(fineLabels[i])) continue;
		const double p[3] = {.5 * ((double)x + .5) - .5, .5 * ((double)y + .5) - .5, .5 * ((double)z + .5) - .5};
		const i64 s[3] = {(i64)p[0], (i64)p[1], (i64)p[2]}; /* truncation, Ops.h:933 */
		const double f[3] = {p[0] - (double)s[0], p[1] - (double)s[1], p[2] - (double)s[2]};
		double v[2][2][2];
		for (int dz = 0; dz < 2; ++dz)
		    for (int dy = 0; dy < 2; ++dy)
			for (int dx = 0; dx < 2; ++dx)
			{
			    /* clamped read like the reference's probe at the grid border (never reached for active cells) */
			    i64 cx = s[0] + dx, cy = s[1] + dy, cz = s[2] + dz;
			    cx = cx < 0 ? 0 : (cx >= cres[0] ? cres[0] - 1 : cx);
			    cy = cy < 0 ? 0 : (cy >= cres[1] ? cres[1] - 1 : cy);
			    cz = cz < 0 ? 0 : (cz >= cres[2] ? cres[2] - 1 : cz);
			    v[dx][dy][dz] = coarse[lin(cres, cx, cy, cz)];
			}
		const double lo = lerp1(lerp1(v[0][0][0], v[1][0][0], f[0]), lerp1(v[0][1][0], v[1][1][0], f[0]), f[1]);
		const double hi = lerp1(lerp1(v[0][0][1], v[1][0][1], f[0]), lerp1(v[0][1][1], v[1][1][1], f[0]), f[1]);
		fine[i] = fine[i] + 4. * lerp1(lo, hi, f[2]);
	    }
}

/* ------------------------------------------------------------------ BLAS-1 (tile-ordered sums, Ops.h:1020-1326) */

typedef double (*tile_fn)(const double *a, const double *b, i64 i);
static inline double fn_dot(const double *a, const double *b, i64 i) { return a[i] * b[i]; }

static double tiled_sum(const double *a, const double *b, const int *labels, const i64 res[3], int isMax)
{
    const i64 tr[3] = {(res[0] + TILE - 1) / TILE, (res[1] + TILE - 1) / TILE, (res[2] + TILE - 1) / TILE};
    const i64 nt = tr[0] * tr[1] * tr[2];
    double *partial = (double *)calloc((size_t)nt, sizeof(double));
#pragma omp parallel for schedule(dynamic, 16)
    for (i64 t = 0; t < nt; ++t)
    {
	const i64 tx = t % tr[0], ty = (t / tr[0]) % tr[1], tz = t / (tr[0] * tr[1]);
	double local = 0;
	for (i64 z = tz * TILE; z < (tz + 1) * TILE && z < res[2]; ++z)
	    for (i64 y = ty * TILE; y < (ty + 1) * TILE && y < res[1]; ++y)
		for (i64 x = tx * TILE; x < (tx + 1) * TILE && x < res[0]; ++x)
		{
		    const i64 i = lin(res, x, y, z);
		    if (!is_active(labels[i])) continue;
		    if (isMax) local = a[i] > local ? a[i] : local; /* max(v, 0), not max|v|: Ops.h:1303-1312 */
		    else local += fn_dot(a, b, i);
		}
	partial[t] = local;
    }
    double acc = 0;
    for (i64 t = 0; t < nt; ++t)
    {
	if (isMax) acc = partial[t] > acc ? partial[t] : acc;
	else acc += partial[t];
    }
    free(partial);
    return acc;
}

double orc_dot(const double *a, const double *b, const int *labels, const i64 res[3]) { return tiled_sum(a, b, labels, res, 0); }
double orc_norm2(const double *a, const int *labels, const i64 res[3]) { return tiled_sum(a, a, labels, res, 0); }
double orc_inf_norm(const double *a, const int *labels, const i64 res[3]) { return tiled_sum(a, a, labels, res, 1); }

/* Ops.h:1087-1137 */
void orc_axpy(double *dst, const double *src, double s, const int *labels, const i64 res[3])
{
    const i64 n = cells_of(res);
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i)
	if (is_active(labels[i])) dst[i] = dst[i] + s * src[i];
}
/* Ops.h:1139-1195 (dst may alias a or v) */
void orc_add_scaled(double *dst, const double *a, const double *v, double s, const int *labels, const i64 res[3])
{
    const i64 n = cells_of(res);
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i)
	if (is_active(labels[i])) dst[i] = a[i] + s * v[i];
}
/* Ops.h:974-1018 */
void orc_scale(double *v, double s, const int *labels, const i64 res[3])
{
    const i64 n = cells_of(res);
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i)
	if (is_active(labels[i])) v[i] = s * v[i];
}

/* ------------------------------------------------------------------ multigrid solver (MG.cpp) */

typedef struct
{
    int levels;
    int allocLevels;
    int useGS;
    i64 (*res)[3];
    int **labels;
    i64 **cells;
    i64 *cellCount;
    double **x, **b, **r;
    double *w[3];
    /* coarsest direct solve */
    i64 nCoarse;
    int *coarseIndex;
    double *L, *D;
    double coarseScale;
} OrcSolver;

static int has_solvable(const int *labels, i64 n)
{
    for (i64 i = 0; i < n; ++i)
	if (is_active(labels[i])) return 1;
    return 0;
}

/* MG.cpp:289-412: number active cells in tile order (x fastest in tile), assemble, factor (dense LDL^T).
 * coarseScale multiplies the matrix: the reference's assembly lambda runs once per UT_ThreadedAlgorithm job
 * without splitting the tile range (MG.cpp:334-389 has no splitByTile), so with J jobs every triplet is
 * emitted J times and setFromTriplets sums them: the factored matrix is J*A. J=1 is the intended algorithm. */
static void build_coarse_solver(OrcSolver *s)
{
    const int lv = s->levels - 1;
    const i64 *res = s->res[lv];
    const int *labels = s->labels[lv];
    const i64 n = cells_of(res);
    s->coarseIndex = (int *)malloc(sizeof(int) * (size_t)n);
    for (i64 i = 0; i < n; ++i) s->coarseIndex[i] = -1;
    const i64 tr[3] = {(res[0] + TILE - 1) / TILE, (res[1] + TILE - 1) / TILE, (res[2] + TILE - 1) / TILE};
    i64 count = 0;
    for (i64 tz = 0; tz < tr[2]; ++tz)
	for (i64 ty = 0; ty < tr[1]; ++ty)
	    for (i64 tx = 0; tx < tr[0]; ++tx)
		for (i64 z = tz * TILE; z < (tz + 1) * TILE && z < res[2]; ++z)
		    for (i64 y = ty * TILE; y < (ty + 1) * TILE && y < res[1]; ++y)
			for (i64 x = tx * TILE; x < (tx + 1) * TILE && x < res[0]; ++x)
			    if (is_active(labels[lin(res, x, y, z)])) s->coarseIndex[lin(res, x, y, z)] = (int)count++;
    s->nCoarse = count;
    const size_t m = (size_t)count;
    double *A = (double *)calloc(m * m + 1, sizeof(double));
    for (i64 z = 0; z < res[2]; ++z)
	for (i64 y = 0; y < res[1]; ++y)
	    for (i64 x = 0; x < res[0]; ++x)
	    {
		const int row = s->coarseIndex[lin(res, x, y, z)];
		if (row < 0) continue;
		double diag = 0;
		for (int axis = 0; axis < 3; ++axis)
		    for (int dir = 0; dir < 2; ++dir)
		    {
			i64 c[3] = {x, y, z};
			c[axis] += dir ? 1 : -1;
			int nl = label_at(labels, res, c[0], c[1], c[2]);
			if (is_active(nl)) { A[(size_t)row * m + (size_t)s->coarseIndex[lin(res, c[0], c[1], c[2])]] += -1 * s->coarseScale; ++diag; }
			else if (nl == DIRICHLET_CELL) ++diag;
		    }
		A[(size_t)row * m + (size_t)row] += diag * s->coarseScale;
	    }
    s->L = (double *)calloc(m * m + 1, sizeof(double));
    s->D = (double *)calloc(m + 1, sizeof(double));
    for (size_t j = 0; j < m; ++j)
    {
	double d = A[j * m + j];
	for (size_t k = 0; k < j; ++k) d -= s->L[j * m + k] * s->L[j * m + k] * s->D[k];
	s->D[j] = d;
	s->L[j * m + j] = 1;
	for (size_t i = j + 1; i < m; ++i)
	{
	    double v = A[i * m + j];
	    for (size_t k = 0; k < j; ++k) v -= s->L[i * m + k] * s->L[j * m + k] * s->D[k];
	    s->L[i * m + j] = v / d;
	}
    }
    free(A);
}

void orc_solver_destroy(OrcSolver *s);

/* MG.cpp:135-418 */
OrcSolver *orc_solver_create(const int *labels, const i64 res[3], const double *w0, const double *w1, const double *w2,
			     int mgLevels, int useGaussSeidel, double coarseScale)
{
    OrcSolver *s = (OrcSolver *)calloc(1, sizeof(OrcSolver));
    s->levels = mgLevels;
    s->allocLevels = mgLevels;
    s->useGS = useGaussSeidel;
    s->coarseScale = coarseScale > 0 ? coarseScale : 1.0;
    s->res = (i64(*)[3])calloc((size_t)mgLevels, sizeof(i64[3]));
    s->labels = (int **)calloc((size_t)mgLevels, sizeof(int *));
    s->cells = (i64 **)calloc((size_t)mgLevels, sizeof(i64 *));
    s->cellCount = (i64 *)calloc((size_t)mgLevels, sizeof(i64));
    s->x = (double **)calloc((size_t)mgLevels, sizeof(double *));
    s->b = (double **)calloc((size_t)mgLevels, sizeof(double *));
    s->r = (double **)calloc((size_t)mgLevels, sizeof(double *));
    const double *ws[3] = {w0, w1, w2};
    for (int a = 0; a < 3; ++a)
    {
	i64 fr[3] = {res[0], res[1], res[2]};
	++fr[a];
	s->w[a] = (double *)malloc(sizeof(double) * (size_t)cells_of(fr));
	memcpy(s->w[a], ws[a], sizeof(double) * (size_t)cells_of(fr));
    }
    for (int a = 0; a < 3; ++a) s->res[0][a] = res[a];
    s->labels[0] = (int *)malloc(sizeof(int) * (size_t)cells_of(res));
    memcpy(s->labels[0], labels, sizeof(int) * (size_t)cells_of(res));
    /* MG.cpp:238-253: coarsen; if a level has no active cell, cap at level-1 (drops one extra level) */
    for (int level = 1; level < s->levels; ++level)
    {
	for (int a = 0; a < 3; ++a) s->res[level][a] = s->res[level - 1][a] / 2;
	s->labels[level] = (int *)malloc(sizeof(int) * (size_t)cells_of(s->res[level]));
	orc_coarsen_labels(s->labels[level - 1], s->res[level - 1], s->labels[level]);
	if (!has_solvable(s->labels[level], cells_of(s->res[level])))
	{
	    s->levels = level - 1;
	    break;
	}
    }
    if (s->levels < 1) { orc_solver_destroy(s); return NULL; }
    for (int level = 0; level < s->levels; ++level)
    {
	const i64 n = cells_of(s->res[level]);
	s->x[level] = (double *)calloc((size_t)n, sizeof(double));
	s->b[level] = (double *)calloc((size_t)n, sizeof(double));
	s->r[level] = (double *)calloc((size_t)n, sizeof(double));
	const i64 cnt = orc_boundary_cells(s->labels[level], s->res[level], 3, NULL, 0);
	s->cells[level] = (i64 *)malloc(sizeof(i64) * 3 * (size_t)(cnt + 1));
	s->cellCount[level] = orc_boundary_cells(s->labels[level], s->res[level], 3, s->cells[level], cnt);
    }
    build_coarse_solver(s);
    return s;
}

void orc_solver_destroy(OrcSolver *s)
{
    if (!s) return;
    for (int l = 0; l < s->allocLevels; ++l)
    {
	free(s->labels[l]); free(s->cells[l]); free(s->x[l]); free(s->b[l]); free(s->r[l]);
    }
    for (int a = 0; a < 3; ++a) free(s->w[a]);
    free(s->res); free(s->labels); free(s->cells); free(s->cellCount); free(s->x); free(s->b); free(s->r);
    free(s->coarseIndex); free(s->L); free(s->D);
    free(s);
}

int orc_solver_levels(const OrcSolver *s) { return s->levels; }
void orc_solver_level_res(const OrcSolver *s, int level, i64 out[3]) { for (int a = 0; a < 3; ++a) out[a] = s->res[level][a]; }
void orc_solver_get_labels(const OrcSolver *s, int level, int *out) { memcpy(out, s->labels[level], sizeof(int) * (size_t)cells_of(s->res[level])); }
i64 orc_solver_boundary_count(const OrcSolver *s, int level) { return s->cellCount[level]; }
void orc_solver_get_boundary_cells(const OrcSolver *s, int level, i64 *xyz) { memcpy(xyz, s->cells[level], sizeof(i64) * 3 * (size_t)s->cellCount[level]); }
i64 orc_solver_coarse_unknowns(const OrcSolver *s) { return s->nCoarse; }

static void smooth_level(OrcSolver *s, int level, double *x, const double *b, int downstroke)
{
    const double *w0 = level == 0 ? s->w[0] : NULL, *w1 = level == 0 ? s->w[1] : NULL, *w2 = level == 0 ? s->w[2] : NULL;
    const i64 *res = s->res[level];
    /* MG.cpp:141-142: 3 band sweeps, width-3 band; interior; 3 band sweeps */
    orc_boundary_jacobi(x, b, s->labels[level], res, s->cells[level], s->cellCount[level], 3, w0, w1, w2);
    if (s->useGS)
    {
	if (downstroke)
	{
	    orc_gauss_seidel(x, b, s->labels[level], res, 1, 1, w0, w1, w2); /* MG.cpp:466-479 */
	    orc_gauss_seidel(x, b, s->labels[level], res, 0, 1, w0, w1, w2);
	}
	else
	{
	    orc_gauss_seidel(x, b, s->labels[level], res, 0, 0, w0, w1, w2); /* MG.cpp:740-751 */
	    orc_gauss_seidel(x, b, s->labels[level], res, 1, 0, w0, w1, w2);
	}
    }
    else
	orc_jacobi(x, b, s->labels[level], res, w0, w1, w2);
    orc_boundary_jacobi(x, b, s->labels[level], res, s->cells[level], s->cellCount[level], 3, w0, w1, w2);
}

/* MG.cpp:420-881 */
void orc_solver_vcycle(OrcSolver *s, double *x, const double *b, int useInitialGuess)
{
    const int L = s->levels;
    if (!useInitialGuess) memset(x, 0, sizeof(double) * (size_t)cells_of(s->res[0]));
    smooth_level(s, 0, x, b, 1);
    if (L == 1) return;
    orc_residual(s->r[0], x, b, s->labels[0], s->res[0], s->w[0], s->w[1], s->w[2]);
    orc_downsample(s->b[1], s->r[0], s->labels[1], s->res[0]);
    for (int level = 1; level < L - 1; ++level)
    {
	memset(s->x[level], 0, sizeof(double) * (size_t)cells_of(s->res[level]));
	smooth_level(s, level, s->x[level], s->b[level], 1);
	orc_residual(s->r[level], s->x[level], s->b[level], s->labels[level], s->res[level], NULL, NULL, NULL);
	orc_downsample(s->b[level + 1], s->r[level], s->labels[level + 1], s->res[level]);
    }
    {
	/* MG.cpp:669-692: gather, L D L^T solve, scatter */
	const int lv = L - 1;
	const i64 n = cells_of(s->res[lv]);
	const size_t m = (size_t)s->nCoarse;
	double *v = (double *)calloc(m + 1, sizeof(double));
	for (i64 i = 0; i < n; ++i)
	    if (s->coarseIndex[i] >= 0) v[s->coarseIndex[i]] = s->b[lv][i];
	for (size_t i = 0; i < m; ++i)
	{
	    double acc = v[i];
	    for (size_t k = 0; k < i; ++k) acc -= s->L[i * m + k] * v[k];
	    v[i] = acc;
	}
	for (size_t i = 0; i < m; ++i) v[i] /= s->D[i];
	for (size_t ii = m; ii-- > 0;)
	{
	    double acc = v[ii];
	    for (size_t k = ii + 1; k < m; ++k) acc -= s->L[k * m + ii] * v[k];
	    v[ii] = acc;
	}
	for (i64 i = 0; i < n; ++i)
	    if (s->coarseIndex[i] >= 0) s->x[lv][i] = v[s->coarseIndex[i]];
	free(v);
    }
    for (int level = L - 2; level >= 1; --level)
    {
	orc_upsample_add(s->x[level], s->x[level + 1], s->labels[level], s->res[level]);
	smooth_level(s, level, s->x[level], s->b[level], 0);
    }
    orc_upsample_add(x, s->x[1], s->labels[0], s->res[0]);
    smooth_level(s, 0, x, b, 0);
}

/* CG.h:11-207 with the operators wired as Test.cpp:746-832 does. Returns the iteration index the reference
 * prints (CG.h:198), or -1 on the two early-outs (CG.h:35-40, :60-64). history[k] is the value CG.h:159 prints. */
static int pcg_impl(OrcSolver *s, double *x, const double *b, double tol, int maxIt, int precond, double *history, int histCap, int *histCount);
int orc_pcg(OrcSolver *s, double *x, const double *b, double tol, int maxIt, double *history, int histCap, int *histCount)
{
    return pcg_impl(s, x, b, tol, maxIt, 1, history, histCap, histCount);
}
/* The node's other mode (GFS.cpp:485-618): the same CG driver with the diagonal preconditioner, 1/6 on INTERIOR cells and
 * 1/(sum of the six face weights) on BOUNDARY cells (GFS.cpp:520-548); destination = source * that (GFS.cpp:598). */
int orc_pcg_diag(OrcSolver *s, double *x, const double *b, double tol, int maxIt, double *history, int histCap, int *histCount)
{
    return pcg_impl(s, x, b, tol, maxIt, 2, history, histCap, histCount);
}
void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
static void precondition(OrcSolver *s, int precond, const double *dinv, double *dst, const double *src)
{
    if (precond == 1) { orc_solver_vcycle(s, dst, src, 0); return; }
    const i64 n = cells_of(s->res[0]);
    const int *labels = s->labels[0];
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i)
	if (is_active(labels[i])) dst[i] = src[i] * dinv[i];
}
static int pcg_impl(OrcSolver *s, double *x, const double *b, double tol, int maxIt, int precond, double *history, int histCap, int *histCount)
{
    const i64 *res = s->res[0];
    const int *labels = s->labels[0];
    const i64 n = cells_of(res);
    *histCount = 0;
    double *dinv = NULL;
    if (precond == 2)
    {
	const double *const w[3] = {s->w[0], s->w[1], s->w[2]};
	dinv = (double *)calloc((size_t)n, sizeof(double));
	for (i64 z = 0; z < res[2]; ++z)
	    for (i64 y = 0; y < res[1]; ++y)
		for (i64 xx = 0; xx < res[0]; ++xx)
		{
		    const i64 i = lin(res, xx, y, z);
		    if (labels[i] == INTERIOR_CELL) dinv[i] = 1. / 6.;
		    else if (labels[i] == BOUNDARY_CELL)
		    {
			double diagonal = 0;
			for (int axis = 0; axis < 3; ++axis)
			    for (int dir = 0; dir < 2; ++dir) diagonal += face_weight(w, res, xx, y, z, axis, dir);
			dinv[i] = 1. / diagonal;
		    }
		}
    }
    const double rhsNorm2 = orc_norm2(b, labels, res);
    if (rhsNorm2 == 0) return -1;
    double *r = (double *)calloc((size_t)n, sizeof(double));
    double *p = (double *)calloc((size_t)n, sizeof(double));
    double *z = (double *)calloc((size_t)n, sizeof(double));
    double *t = (double *)calloc((size_t)n, sizeof(double));
    orc_apply(r, x, labels, res, s->w[0], s->w[1], s->w[2]);
    orc_add_scaled(r, b, r, -1, labels, res);
    double rNorm2 = orc_norm2(r, labels, res);
    const double threshold = tol * tol * rhsNorm2;
    int iteration = -1;
    if (!(rNorm2 < threshold))
    {
	precondition(s, precond, dinv, p, r);
	double absNew = orc_dot(p, r, labels, res);
	for (iteration = 0; iteration < maxIt; ++iteration)
	{
	    orc_apply(t, p, labels, res, s->w[0], s->w[1], s->w[2]);
	    const double alpha = absNew / orc_dot(p, t, labels, res);
	    orc_axpy(x, p, alpha, labels, res);
	    orc_axpy(r, t, -alpha, labels, res);
	    rNorm2 = orc_norm2(r, labels, res);
	    if (*histCount < histCap) history[(*histCount)++] = sqrt(rNorm2 / rhsNorm2);
	    if (rNorm2 < threshold) break;
	    precondition(s, precond, dinv, z, r);
	    const double absOld = absNew;
	    absNew = orc_dot(z, r, labels, res);
	    const double beta = absNew / absOld;
	    orc_add_scaled(p, z, p, beta, labels, res);
	}
    }
    free(r); free(p); free(z); free(t); free(dinv);
    return iteration;
}

/* =====================================================================================================================
 * The steps either side of the path: HDK_GeometricFreeSurfacePressureSolver.cpp builds the solver's
 * inputs from the simulation's fields and applies its output to them.  PINNED: that file and HDK_Utilities.cpp compile
 * unmodified over oracle/shim/hdk_node_shim.h (SIM_RawField, SIM_RawIndexField, SIM_VectorField, GAS_SubSolver stand-ins) into
 * oracle/_ref, and tests/test_node_reference.py / tests/test_oracle_vs_reference.py hold every function below to them bit for
 * bit (and the chain of them to the node's whole solveGasSubclass).  The restatement is function by function, on plain arrays: cell fields [rz][ry][rx] x-fastest, the face field of axis a
 * has one more entry along a, SIM_RawField values are fpreal32 (float) and the arithmetic is SolveReal = double
 * (GFS.h:18-19).  Assumption where the HDK's own return types matter (encoded in the shim, not verifiable offline):
 * SIM::FieldUtils::getFieldValue returns the field's fpreal32, so pressure(forward) - pressure(backward) (GFS.cpp:1095) is a
 * float subtraction.
 * Material labels: HDK_Utilities.h:17 { SOLID_CELL = 0, LIQUID_CELL = 1, AIR_CELL = 2 }; VALID_FACE = 1 (HDK_Utilities.h:21).
 * ===================================================================================================================== */
enum { MAT_SOLID = 0, MAT_LIQUID = 1, MAT_AIR = 2 };

/* HDK_Utilities.h:25-42 */
static inline double ghost_fluid_weight(double phi0, double phi1)
{
    double theta = 0;
    if (phi0 < 0)
    {
	if (phi1 < 0) theta = 1;
	else if (phi1 >= 0) theta = phi0 / (phi0 - phi1);
    }
    else if (phi1 < 0) theta = phi1 / (phi1 - phi0);
    return theta;
}
static inline double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* HDK_Utilities.cpp:5-45 isCellLiquid.  solidAtCentres = solidSurface.getValue(indexToPos(cell)) (:21-24): the solid SDF sampled at
 * the surface field's cell centres -- an HDK interpolation that is not restated; the caller supplies the samples. */
static int is_cell_liquid(const float *liquidSurface, const float *solidAtCentres, const float *const cutCell[3], const i64 res[3], i64 x, i64 y, i64 z)
{
    if (liquidSurface[lin(res, x, y, z)] <= 0.) return 1;
    if (solidAtCentres[lin(res, x, y, z)] >= 0)
    {
	for (int axis = 0; axis < 3; ++axis)
	    for (int direction = 0; direction < 2; ++direction)
	    {
		i64 fr[3] = {res[0], res[1], res[2]};
		++fr[axis];
		i64 fc[3] = {x, y, z};
		fc[axis] += direction;  /* cellToFaceMap */
		if (cutCell[axis][lin(fr, fc[0], fc[1], fc[2])] > 0)
		{
		    i64 ac[3] = {x, y, z};
		    ac[axis] += direction == 0 ? -1 : 1;  /* cellToCellMap */
		    if (ac[axis] < 0 || ac[axis] >= res[axis]) continue;
		    if (liquidSurface[lin(res, ac[0], ac[1], ac[2])] <= 0) return 1;
		}
	    }
    }
    return 0;
}

/* HDK_Utilities.cpp:87-148 buildMaterialCellLabels: SOLID_CELL everywhere (:99), then every cell with an open face (a cut-cell
 * weight > 0 on one of its six faces, :122-133) becomes LIQUID_CELL or AIR_CELL by isCellLiquid (:135-141) */
void orc_build_material_labels(const float *liquidSurface, const float *solidAtCentres, const float *const cutCell[3], const i64 res[3], int *material)
{
    for (i64 z = 0; z < res[2]; ++z)
	for (i64 y = 0; y < res[1]; ++y)
	    for (i64 x = 0; x < res[0]; ++x)
	    {
		int isInFluid = 0;
		for (int axis = 0; axis < 3; ++axis)
		    for (int direction = 0; direction < 2; ++direction)
		    {
			i64 fr[3] = {res[0], res[1], res[2]};
			++fr[axis];
			i64 fc[3] = {x, y, z};
			fc[axis] += direction;
			if (cutCell[axis][lin(fr, fc[0], fc[1], fc[2])] > 0) isInFluid = 1;
		    }
		int label = MAT_SOLID;
		if (isInFluid) label = is_cell_liquid(liquidSurface, solidAtCentres, cutCell, res, x, y, z) ? MAT_LIQUID : MAT_AIR;
		material[lin(res, x, y, z)] = label;
	    }
}

/* GFS.cpp:717-744 buildValidFaces for one axis: INVALID_FACE everywhere (:724), then classifyValidFaces (HDK_Utilities.h:137-189):
 * VALID_FACE where the cut-cell weight is > 0 (:173), both cells of the face are in range (:180) and one of them is LIQUID (:182-183).
 * (The tile bookkeeping in between, findOccupiedFaceTiles / uncompressTiles, only uncompresses the tiles that can hold such a face.) */
void orc_build_valid_faces(const int *material, const float *cutCell, const i64 res[3], int axis, float *validFaces)
{
    i64 fr[3] = {res[0], res[1], res[2]};
    ++fr[axis];
    for (i64 z = 0; z < fr[2]; ++z)
	for (i64 y = 0; y < fr[1]; ++y)
	    for (i64 x = 0; x < fr[0]; ++x)
	    {
		const i64 f = lin(fr, x, y, z);
		validFaces[f] = 0.f;
		if (!(cutCell[f] > 0)) continue;
		i64 b[3] = {x, y, z}, fw[3] = {x, y, z};
		--b[axis];  /* faceToCellMap(face, axis, 0) */
		if (b[axis] >= 0 && fw[axis] < res[axis])
		    if (material[lin(res, b[0], b[1], b[2])] == MAT_LIQUID || material[lin(res, fw[0], fw[1], fw[2])] == MAT_LIQUID) validFaces[f] = 1.f;
	    }
}

/* GFS.cpp:746-793 buildMGDomainLabels: LIQUID -> INTERIOR, AIR -> DIRICHLET, everything else stays EXTERIOR (GFS.cpp:309) */
void orc_build_domain_labels(const int *material, const i64 res[3], int *labels)
{
    const i64 n = cells_of(res);
    for (i64 i = 0; i < n; ++i)
	labels[i] = material[i] == MAT_LIQUID ? INTERIOR_CELL : (material[i] == MAT_AIR ? DIRICHLET_CELL : EXTERIOR_CELL);
}

/* GFS.cpp:796-865 buildMGBoundaryWeights for one axis: weights[face] = 0 (GFS.cpp:322), then on VALID faces the cut-cell
 * weight, divided by the clamped ghost-fluid theta on a liquid/air face */
void orc_build_boundary_weights(const float *cutCell, const float *liquidSurface, const float *validFaces, const int *domainLabels, const i64 res[3],
				int axis, double *weights)
{
    i64 fr[3] = {res[0], res[1], res[2]};
    ++fr[axis];
    for (i64 z = 0; z < fr[2]; ++z)
	for (i64 y = 0; y < fr[1]; ++y)
	    for (i64 x = 0; x < fr[0]; ++x)
	    {
		const i64 f = lin(fr, x, y, z);
		weights[f] = 0;
		if (validFaces[f] != 1.0f) continue;
		double weight = cutCell[f];
		i64 b[3] = {x, y, z}, fw[3] = {x, y, z};
		--b[axis];  /* faceToCellMap(face, axis, 0) */
		const int bl = label_at(domainLabels, res, b[0], b[1], b[2]), fl = label_at(domainLabels, res, fw[0], fw[1], fw[2]);
		if ((bl == INTERIOR_CELL && fl == DIRICHLET_CELL) || (bl == DIRICHLET_CELL && fl == INTERIOR_CELL))
		{
		    const double phi0 = liquidSurface[lin(res, b[0], b[1], b[2])], phi1 = liquidSurface[lin(res, fw[0], fw[1], fw[2])];
		    double theta = ghost_fluid_weight(phi0, phi1);
		    theta = clampd(theta, .01, 1.);
		    weight /= theta;
		}
		weights[f] = weight;
	    }
}

/* GFS.cpp:868-943 buildRHS: cut-cell divergence of every LIQUID cell into the expanded rhs grid.  solidVelocity (nullable)
 * holds the solid's velocity component already sampled at every face centre of the axis (the reference samples a
 * SIM_VectorField at the face position, GFS.cpp:918-921: an HDK interpolation that is not restated). */
void orc_build_rhs(const int *material, const float *const velocity[3], const float *const cutCell[3], const float *const solidVelocity[3], const i64 res[3],
		   const i64 expRes[3], const i64 offset[3], double *rhs)
{
    for (i64 z = 0; z < res[2]; ++z)
	for (i64 y = 0; y < res[1]; ++y)
	    for (i64 x = 0; x < res[0]; ++x)
	    {
		if (material[lin(res, x, y, z)] != MAT_LIQUID) continue;
		double divergence = 0;
		for (int axis = 0; axis < 3; ++axis)
		    for (int direction = 0; direction < 2; ++direction)
		    {
			i64 fr[3] = {res[0], res[1], res[2]};
			++fr[axis];
			i64 fc[3] = {x, y, z};
			fc[axis] += direction;  /* cellToFaceMap */
			const i64 f = lin(fr, fc[0], fc[1], fc[2]);
			const double sign = (direction == 0) ? 1. : -1.;
			const double weight = cutCell[axis][f];
			if (weight > 0) divergence += sign * weight * velocity[axis][f];
			if (solidVelocity && weight < 1) divergence += sign * (1. - weight) * solidVelocity[axis][f];
		    }
		rhs[lin(expRes, x + offset[0], y + offset[1], z + offset[2])] = divergence;
	    }
}

/* GFS.cpp:946-997 applyOldPressure: the warm start */
void orc_apply_old_pressure(const float *pressure, const int *material, const i64 res[3], const i64 expRes[3], const i64 offset[3], double *solution)
{
    for (i64 z = 0; z < res[2]; ++z)
	for (i64 y = 0; y < res[1]; ++y)
	    for (i64 x = 0; x < res[0]; ++x)
		if (material[lin(res, x, y, z)] == MAT_LIQUID) solution[lin(expRes, x + offset[0], y + offset[1], z + offset[2])] = pressure[lin(res, x, y, z)];
}

/* GFS.cpp:1000-1047 applySolutionToPressure (the pressure field is fpreal32: the store rounds) */
void orc_apply_solution_to_pressure(float *pressure, const int *material, const double *solution, const i64 res[3], const i64 expRes[3], const i64 offset[3])
{
    for (i64 z = 0; z < res[2]; ++z)
	for (i64 y = 0; y < res[1]; ++y)
	    for (i64 x = 0; x < res[0]; ++x)
		if (material[lin(res, x, y, z)] == MAT_LIQUID) pressure[lin(res, x, y, z)] = (float)solution[lin(expRes, x + offset[0], y + offset[1], z + offset[2])];
}

/* GFS.cpp:1050-1131 applyPressureGradient for one axis */
void orc_apply_pressure_gradient(float *velocity, const float *cutCell, const float *liquidSurface, const float *pressure, const float *validFaces,
				 const int *material, const i64 res[3], int axis)
{
    (void)cutCell; /* only asserted > 0 (GFS.cpp:1093) */
    i64 fr[3] = {res[0], res[1], res[2]};
    ++fr[axis];
    for (i64 z = 0; z < fr[2]; ++z)
	for (i64 y = 0; y < fr[1]; ++y)
	    for (i64 x = 0; x < fr[0]; ++x)
	    {
		const i64 f = lin(fr, x, y, z);
		if (validFaces[f] != 1.0f) continue;
		i64 b[3] = {x, y, z}, fw[3] = {x, y, z};
		--b[axis];
		if (b[axis] < 0 || fw[axis] >= res[axis]) continue;
		const i64 bi = lin(res, b[0], b[1], b[2]), fi = lin(res, fw[0], fw[1], fw[2]);
		const int bm = material[bi], fm = material[fi];
		double gradient = (float)(pressure[fi] - pressure[bi]);
		if (bm != MAT_LIQUID || fm != MAT_LIQUID)
		{
		    double theta = ghost_fluid_weight(liquidSurface[bi], liquidSurface[fi]);
		    theta = clampd(theta, .01, 1.);
		    gradient /= theta;
		}
		velocity[f] = (float)(velocity[f] - gradient);
	    }
}
