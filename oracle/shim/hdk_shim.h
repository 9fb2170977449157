// hdk_shim.h -- TEST INFRASTRUCTURE ONLY (oracle build).
//
// Minimal stand-in for the slice of the Houdini 18 HDK that the reference's
// hot-path sources use (HDK_GeometricMultigridOperators.{h,cpp},
// HDK_GeometricMultigridPoissonSolver.{h,cpp}, HDK_GeometricCGPoissonSolver.h, HDK_Utilities.{h,cpp}),
// so that those files compile UNMODIFIED from /root/reference/Source into
// oracle/_ref/.  The symbol list follows SURVEY.md appendix B.
//
// Semantics encoded here (taken from the public HDK 18 documentation, not
// verifiable offline):
//   * UT_VoxelArray stores 16^3 tiles, x-fastest inside a tile, linear tile
//     index x-fastest; a tile is either "constant" (one value) or a dense block.
//   * operator() reads clamp out-of-range indices to the border.
//   * iterators visit tiles in linear order and voxels x-fastest in a tile.
//   * UT_Vector3I is 3 x int64.
//   * UTparallelFor* run the body over sub-ranges (OpenMP here, TBB there).
//   * UT_ThreadedAlgorithm::run runs the body once per job.  The job count is
//     GMG_SHIM_JOBS (default 1) -- see DESIGN.md "coarse-matrix job-count quirk".
//
// Nothing in the shipped product includes this file.
#ifndef GMG_ORACLE_HDK_SHIM_H
#define GMG_ORACLE_HDK_SHIM_H

#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <iostream>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

// ---------------------------------------------------------------- SYS types
using fpreal = double;
using fpreal32 = float;
using fpreal64 = double;
using exint = int64_t;

#define SYS_FORCE_INLINE inline __attribute__((always_inline))

template <typename T> static inline T SYSsqrt(T v) { return std::sqrt(v); }
template <typename T> static inline T SYSsin(T v) { return std::sin(v); }
template <typename T> static inline T SYSmax(T a, T b) { return a > b ? a : b; }
template <typename T> static inline T SYSmin(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T SYSclamp(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---------------------------------------------------------------- UT_Vector3T
template <typename T>
class UT_Vector3T
{
public:
    UT_Vector3T() : v{0, 0, 0} {}
    UT_Vector3T(T x, T y, T z) : v{x, y, z} {}
    explicit UT_Vector3T(T s) : v{s, s, s} {}
    template <typename S>
    explicit UT_Vector3T(const UT_Vector3T<S> &o) : v{T(o[0]), T(o[1]), T(o[2])} {}

    T &operator[](int i) { return v[i]; }
    const T &operator[](int i) const { return v[i]; }
    T &operator()(int i) { return v[i]; }
    const T &operator()(int i) const { return v[i]; }
    T x() const { return v[0]; }
    T y() const { return v[1]; }
    T z() const { return v[2]; }

    UT_Vector3T &operator+=(const UT_Vector3T &o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
    UT_Vector3T &operator-=(const UT_Vector3T &o) { v[0] -= o.v[0]; v[1] -= o.v[1]; v[2] -= o.v[2]; return *this; }

    bool operator==(const UT_Vector3T &o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
    bool operator!=(const UT_Vector3T &o) const { return !(*this == o); }
    T maxComponent() const { return std::max(v[0], std::max(v[1], v[2])); }

private:
    T v[3];
};

template <typename T> inline UT_Vector3T<T> operator+(const UT_Vector3T<T> &a, const UT_Vector3T<T> &b) { return UT_Vector3T<T>(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
template <typename T> inline UT_Vector3T<T> operator-(const UT_Vector3T<T> &a, const UT_Vector3T<T> &b) { return UT_Vector3T<T>(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
template <typename T> inline UT_Vector3T<T> operator*(const UT_Vector3T<T> &a, const UT_Vector3T<T> &b) { return UT_Vector3T<T>(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
template <typename T> inline UT_Vector3T<T> operator/(const UT_Vector3T<T> &a, const UT_Vector3T<T> &b) { return UT_Vector3T<T>(a[0] / b[0], a[1] / b[1], a[2] / b[2]); }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline UT_Vector3T<T> operator*(S s, const UT_Vector3T<T> &a) { return UT_Vector3T<T>(T(s * a[0]), T(s * a[1]), T(s * a[2])); }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline UT_Vector3T<T> operator*(const UT_Vector3T<T> &a, S s) { return UT_Vector3T<T>(T(a[0] * s), T(a[1] * s), T(a[2] * s)); }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline UT_Vector3T<T> operator/(const UT_Vector3T<T> &a, S s) { return UT_Vector3T<T>(T(a[0] / s), T(a[1] / s), T(a[2] / s)); }

using UT_Vector3I = UT_Vector3T<int64_t>;
using UT_Vector3i = UT_Vector3T<int32_t>;
using UT_Vector3 = UT_Vector3T<fpreal32>;
using UT_Vector3D = UT_Vector3T<fpreal64>;

template <typename T>
inline T distance2(const UT_Vector3T<T> &a, const UT_Vector3T<T> &b)
{
    T d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    return d0 * d0 + d1 * d1 + d2 * d2;
}

// ---------------------------------------------------------------- UT_Array
template <typename T>
class UT_Array
{
public:
    UT_Array() {}
    UT_Array(const UT_Array &o) { copyFrom(o); }
    UT_Array(UT_Array &&o) noexcept : myData(o.myData), mySize(o.mySize), myCap(o.myCap) { o.myData = nullptr; o.mySize = o.myCap = 0; }
    ~UT_Array() { destroy(); }
    UT_Array &operator=(const UT_Array &o) { if (this != &o) { destroy(); copyFrom(o); } return *this; }
    UT_Array &operator=(UT_Array &&o) noexcept
    {
	if (this != &o) { destroy(); myData = o.myData; mySize = o.mySize; myCap = o.myCap; o.myData = nullptr; o.mySize = o.myCap = 0; }
	return *this;
    }

    exint size() const { return mySize; }
    exint entries() const { return mySize; }
    exint capacity() const { return myCap; }

    void setSize(exint n)
    {
	if (n > myCap) grow(n);
	for (exint i = mySize; i < n; ++i) new (&myData[i]) T();
	for (exint i = n; i < mySize; ++i) myData[i].~T();
	mySize = n;
    }
    void bumpCapacity(exint n) { if (n > myCap) grow(n); }
    void setCapacity(exint n) { if (n > myCap) grow(n); }
    void clear() { for (exint i = 0; i < mySize; ++i) myData[i].~T(); mySize = 0; }
    void constant(const T &v) { for (exint i = 0; i < mySize; ++i) myData[i] = v; }
    exint append(const T &v)
    {
	if (mySize == myCap) { T tmp(v); grow(myCap ? 2 * myCap : 16); new (&myData[mySize]) T(std::move(tmp)); }
	else new (&myData[mySize]) T(v);
	return mySize++;
    }
    void concat(const UT_Array &o)
    {
	if (mySize + o.mySize > myCap) grow(mySize + o.mySize);
	for (exint i = 0; i < o.mySize; ++i) new (&myData[mySize + i]) T(o.myData[i]);
	mySize += o.mySize;
    }

    T &operator[](exint i) { return myData[i]; }
    const T &operator[](exint i) const { return myData[i]; }
    T &operator()(exint i) { return myData[i]; }
    const T &operator()(exint i) const { return myData[i]; }
    T *begin() { return myData; }
    T *end() { return myData + mySize; }
    const T *begin() const { return myData; }
    const T *end() const { return myData + mySize; }
    T *data() { return myData; }
    const T *data() const { return myData; }

private:
    void grow(exint n)
    {
	T *nd = static_cast<T *>(::operator new(sizeof(T) * size_t(n)));
	for (exint i = 0; i < mySize; ++i) { new (&nd[i]) T(std::move(myData[i])); myData[i].~T(); }
	::operator delete(myData);
	myData = nd;
	myCap = n;
    }
    void destroy()
    {
	for (exint i = 0; i < mySize; ++i) myData[i].~T();
	::operator delete(myData);
	myData = nullptr;
	mySize = myCap = 0;
    }
    void copyFrom(const UT_Array &o)
    {
	myData = nullptr; mySize = myCap = 0;
	if (o.mySize) { grow(o.mySize); for (exint i = 0; i < o.mySize; ++i) new (&myData[i]) T(o.myData[i]); mySize = o.mySize; }
    }

    T *myData = nullptr;
    exint mySize = 0, myCap = 0;
};

// ---------------------------------------------------------------- interrupt / timers / threads
class UT_Interrupt
{
public:
    bool opInterrupt(int = -1) { return false; }
};
inline UT_Interrupt *UTgetInterrupt() { static UT_Interrupt boss; return &boss; }

class UT_StopWatch
{
public:
    void start() { myStart = std::chrono::steady_clock::now(); myRunning = true; }
    double stop()
    {
	if (myRunning) { myElapsed += std::chrono::duration<double>(std::chrono::steady_clock::now() - myStart).count(); myRunning = false; }
	return myElapsed;
    }
    void clear() { myElapsed = 0; myRunning = false; }
    double lap() const { return myElapsed; }

private:
    std::chrono::steady_clock::time_point myStart;
    double myElapsed = 0;
    bool myRunning = false;
};

inline int gmgShimJobCount()
{
    static int jobs = [] { const char *e = std::getenv("GMG_SHIM_JOBS"); int j = e ? std::atoi(e) : 1; return j > 0 ? j : 1; }();
    return jobs;
}

class UT_Thread
{
public:
    static int getNumProcessors() { return gmgShimJobCount(); }
};

class UT_JobInfo
{
public:
    UT_JobInfo(int job, int numJobs) : myJob(job), myNumJobs(numJobs) {}
    int job() const { return myJob; }
    int numJobs() const { return myNumJobs; }
    void divideWork(exint units, exint &start, exint &end) const
    {
	exint per = (units + myNumJobs - 1) / myNumJobs;
	start = std::min<exint>(units, per * myJob);
	end = std::min<exint>(units, start + per);
    }

private:
    int myJob, myNumJobs;
};

class UT_ThreadedAlgorithm
{
public:
    template <typename Body>
    void run(const Body &body)
    {
	const int jobs = UT_Thread::getNumProcessors();
#pragma omp parallel for schedule(static, 1) if (jobs > 1)
	for (int j = 0; j < jobs; ++j)
	{
	    UT_JobInfo info(j, jobs);
	    body(info);
	}
    }
};

template <typename T>
class UT_BlockedRange
{
public:
    UT_BlockedRange(T b, T e, size_t = 1) : myBegin(b), myEnd(e) {}
    T begin() const { return myBegin; }
    T end() const { return myEnd; }

private:
    T myBegin, myEnd;
};

template <typename IntT, typename Body>
inline void gmgShimParallelRange(IntT begin, IntT end, IntT grain, const Body &body)
{
    const IntT n = end - begin;
    if (n <= 0) return;
    const IntT chunks = (n + grain - 1) / grain;
#pragma omp parallel for schedule(dynamic, 1) if (chunks > 1)
    for (IntT c = 0; c < chunks; ++c)
    {
	IntT b = begin + c * grain;
	IntT e = std::min<IntT>(end, b + grain);
	body(UT_BlockedRange<IntT>(b, e));
    }
}

template <typename IntT, typename Body>
inline void UTparallelForEachNumber(IntT nitems, const Body &body)
{
    gmgShimParallelRange<IntT>(0, nitems, 8, body);
}
template <typename IntT, typename Body>
inline void UTparallelFor(const UT_BlockedRange<IntT> &range, const Body &body)
{
    gmgShimParallelRange<IntT>(range.begin(), range.end(), 64, body);
}
template <typename IntT, typename Body>
inline void UTparallelForLightItems(const UT_BlockedRange<IntT> &range, const Body &body)
{
    gmgShimParallelRange<IntT>(range.begin(), range.end(), 2048, body);
}
template <typename It, typename Cmp>
inline void UTparallelSort(It b, It e, const Cmp &cmp) { std::sort(b, e, cmp); }

// ---------------------------------------------------------------- UT_VoxelArray
constexpr int GMG_TILEBITS = 4;
constexpr int GMG_TILESIZE = 1 << GMG_TILEBITS;
constexpr int GMG_TILEMASK = GMG_TILESIZE - 1;

template <typename T>
class UT_VoxelTile
{
public:
    UT_VoxelTile() {}
    UT_VoxelTile(const UT_VoxelTile &o) { copyFrom(o); }
    UT_VoxelTile &operator=(const UT_VoxelTile &o) { if (this != &o) { delete[] myData; myData = nullptr; copyFrom(o); } return *this; }
    ~UT_VoxelTile() { delete[] myData; }

    bool isConstant() const { return myData == nullptr; }
    int xres() const { return myRes[0]; }
    int yres() const { return myRes[1]; }
    int zres() const { return myRes[2]; }
    int numVoxels() const { return myRes[0] * myRes[1] * myRes[2]; }

    void setRes(int x, int y, int z) { myRes[0] = x; myRes[1] = y; myRes[2] = z; }
    void makeConstant(T v) { delete[] myData; myData = nullptr; myConst = v; }
    void uncompress()
    {
	if (myData) return;
	const int n = numVoxels();
	T *d = new T[n];
	for (int i = 0; i < n; ++i) d[i] = myConst;
	myData = d;
    }
    bool tryCompress()
    {
	if (!myData) return true;
	const int n = numVoxels();
	const T v = myData[0];
	for (int i = 1; i < n; ++i) if (!(myData[i] == v)) return false;
	makeConstant(v);
	return true;
    }
    T get(int lx, int ly, int lz) const { return myData ? myData[(lz * myRes[1] + ly) * myRes[0] + lx] : myConst; }
    void set(int lx, int ly, int lz, T v)
    {
	if (!myData) { if (v == myConst) return; uncompress(); }
	myData[(lz * myRes[1] + ly) * myRes[0] + lx] = v;
    }
    T operator()(int lx, int ly, int lz) const { return get(lx, ly, lz); } // HDK: UT_VoxelTile::operator()(x, y, z)
    T constantValue() const { return myConst; }
    T *rawData() { return myData; }
    const T *rawData() const { return myData; }

private:
    void copyFrom(const UT_VoxelTile &o)
    {
	myRes[0] = o.myRes[0]; myRes[1] = o.myRes[1]; myRes[2] = o.myRes[2];
	myConst = o.myConst;
	if (o.myData) { const int n = numVoxels(); myData = new T[n]; std::memcpy(myData, o.myData, sizeof(T) * n); }
    }
    T *myData = nullptr;
    T myConst = T(0);
    int myRes[3] = {0, 0, 0};
};

template <typename T>
class UT_VoxelArray
{
public:
    UT_VoxelArray() {}

    void size(int x, int y, int z)
    {
	myRes[0] = x; myRes[1] = y; myRes[2] = z;
	for (int a = 0; a < 3; ++a) myTileRes[a] = (myRes[a] + GMG_TILESIZE - 1) >> GMG_TILEBITS;
	myTiles.assign(size_t(myTileRes[0]) * myTileRes[1] * myTileRes[2], UT_VoxelTile<T>());
	for (int tz = 0; tz < myTileRes[2]; ++tz)
	    for (int ty = 0; ty < myTileRes[1]; ++ty)
		for (int tx = 0; tx < myTileRes[0]; ++tx)
		    myTiles[(size_t(tz) * myTileRes[1] + ty) * myTileRes[0] + tx].setRes(std::min(GMG_TILESIZE, x - tx * GMG_TILESIZE),
											  std::min(GMG_TILESIZE, y - ty * GMG_TILESIZE),
											  std::min(GMG_TILESIZE, z - tz * GMG_TILESIZE));
    }
    void constant(T v) { for (auto &t : myTiles) t.makeConstant(v); }
    // every tile constant-compressed with one and the same value
    bool isConstant(T *cval = nullptr) const
    {
	for (const auto &t : myTiles)
	    if (!t.isConstant() || !(t.constantValue() == myTiles[0].constantValue())) return false;
	if (cval && !myTiles.empty()) *cval = myTiles[0].constantValue();
	return true;
    }

    UT_Vector3I getVoxelRes() const { return UT_Vector3I(myRes[0], myRes[1], myRes[2]); }
    int getXRes() const { return myRes[0]; }
    int getYRes() const { return myRes[1]; }
    int getZRes() const { return myRes[2]; }
    int getTileRes(int a) const { return myTileRes[a]; }
    int numTiles() const { return int(myTiles.size()); }

    UT_VoxelTile<T> *getLinearTile(int i) const { return const_cast<UT_VoxelTile<T> *>(&myTiles[i]); }
    int indexToLinearTile(int x, int y, int z) const
    {
	return ((z >> GMG_TILEBITS) * myTileRes[1] + (y >> GMG_TILEBITS)) * myTileRes[0] + (x >> GMG_TILEBITS);
    }
    void linearTileToXYZ(int i, int &x, int &y, int &z) const
    {
	x = i % myTileRes[0];
	i /= myTileRes[0];
	y = i % myTileRes[1];
	z = i / myTileRes[1];
    }

    // Clamped read (HDK default border behaviour for operator()).
    T operator()(int x, int y, int z) const
    {
	x = x < 0 ? 0 : (x >= myRes[0] ? myRes[0] - 1 : x);
	y = y < 0 ? 0 : (y >= myRes[1] ? myRes[1] - 1 : y);
	z = z < 0 ? 0 : (z >= myRes[2] ? myRes[2] - 1 : z);
	return myTiles[indexToLinearTile(x, y, z)].get(x & GMG_TILEMASK, y & GMG_TILEMASK, z & GMG_TILEMASK);
    }
    T operator()(const UT_Vector3I &c) const { return (*this)(int(c[0]), int(c[1]), int(c[2])); }
    T getValue(int x, int y, int z) const { return (*this)(x, y, z); }

    void setValue(int x, int y, int z, T v)
    {
	myTiles[indexToLinearTile(x, y, z)].set(x & GMG_TILEMASK, y & GMG_TILEMASK, z & GMG_TILEMASK, v);
    }
    void setValue(const UT_Vector3I &c, T v) { setValue(int(c[0]), int(c[1]), int(c[2]), v); }

    void collapseAllTiles()
    {
	const int n = numTiles();
#pragma omp parallel for schedule(dynamic, 64)
	for (int i = 0; i < n; ++i) myTiles[i].tryCompress();
    }

private:
    int myRes[3] = {0, 0, 0};
    int myTileRes[3] = {0, 0, 0};
    std::vector<UT_VoxelTile<T>> myTiles;
};

// ---------------------------------------------------------------- iterators
template <typename T> class UT_VoxelTileIterator;

template <typename T>
class UT_VoxelArrayIterator
{
public:
    UT_VoxelArrayIterator() {}
    explicit UT_VoxelArrayIterator(UT_VoxelArray<T> *a) { setArray(a); }

    void setArray(UT_VoxelArray<T> *a) { myArray = a; myTileStart = 0; myTileEnd = a->numTiles(); rewind(); }
    void setConstArray(const UT_VoxelArray<T> *a) { setArray(const_cast<UT_VoxelArray<T> *>(a)); }

    void splitByTile(const UT_JobInfo &info)
    {
	exint s, e;
	info.divideWork(myArray->numTiles(), s, e);
	myTileStart = int(s);
	myTileEnd = int(e);
	rewind();
    }

    void rewind() { myCurTile = myTileStart; enterTile(); }
    bool atEnd() const { return myCurTile >= myTileEnd; }
    void advance()
    {
	++myIdx;
	if (++myLocal[0] < myTileDim[0]) return;
	myLocal[0] = 0;
	if (++myLocal[1] < myTileDim[1]) return;
	myLocal[1] = 0;
	if (++myLocal[2] < myTileDim[2]) return;
	advanceTile();
    }
    void advanceTile() { ++myCurTile; enterTile(); }

    bool isTileConstant() const { return myTile->isConstant(); }
    T getValue() const { return myTile->isConstant() ? myTile->constantValue() : myTile->rawData()[myIdx]; }
    void setValue(T v) { myTile->set(myLocal[0], myLocal[1], myLocal[2], v); }
    void setCompressOnExit(bool) {}  // HDK: re-compress a tile when the iterator leaves it; storage only, values unaffected

    int x() const { return myTileOrigin[0] + myLocal[0]; }
    int y() const { return myTileOrigin[1] + myLocal[1]; }
    int z() const { return myTileOrigin[2] + myLocal[2]; }
    int getLinearTileNum() const { return myCurTile; }
    void getTileVoxels(UT_Vector3I &start, UT_Vector3I &end) const
    {
	start = UT_Vector3I(myTileOrigin[0], myTileOrigin[1], myTileOrigin[2]);
	end = UT_Vector3I(myTileOrigin[0] + myTileDim[0], myTileOrigin[1] + myTileDim[1], myTileOrigin[2] + myTileDim[2]);
    }

    int myTileStart = 0, myTileEnd = 0;

private:
    friend class UT_VoxelTileIterator<T>;
    void enterTile()
    {
	myLocal[0] = myLocal[1] = myLocal[2] = 0;
	myIdx = 0;
	if (myArray && myCurTile < myTileEnd && myCurTile < myArray->numTiles())
	{
	    myTile = myArray->getLinearTile(myCurTile);
	    int tx, ty, tz;
	    myArray->linearTileToXYZ(myCurTile, tx, ty, tz);
	    myTileOrigin[0] = tx << GMG_TILEBITS; myTileOrigin[1] = ty << GMG_TILEBITS; myTileOrigin[2] = tz << GMG_TILEBITS;
	    myTileDim[0] = myTile->xres(); myTileDim[1] = myTile->yres(); myTileDim[2] = myTile->zres();
	}
	else myTile = nullptr;
    }

    UT_VoxelArray<T> *myArray = nullptr;
    UT_VoxelTile<T> *myTile = nullptr;
    int myCurTile = 0;
    int myLocal[3] = {0, 0, 0};
    int myTileOrigin[3] = {0, 0, 0};
    int myTileDim[3] = {0, 0, 0};
    int myIdx = 0;
};

template <typename T>
class UT_VoxelTileIterator
{
public:
    void setTile(const UT_VoxelArrayIterator<T> &vit)
    {
	myTile = vit.myTile;
	for (int a = 0; a < 3; ++a) { myTileOrigin[a] = vit.myTileOrigin[a]; myTileDim[a] = vit.myTileDim[a]; }
	rewind();
    }
    void rewind() { myLocal[0] = myLocal[1] = myLocal[2] = 0; myIdx = 0; myDone = (myTile == nullptr); }
    bool atEnd() const { return myDone; }
    void advance()
    {
	++myIdx;
	if (++myLocal[0] < myTileDim[0]) return;
	myLocal[0] = 0;
	if (++myLocal[1] < myTileDim[1]) return;
	myLocal[1] = 0;
	if (++myLocal[2] < myTileDim[2]) return;
	myDone = true;
    }
    T getValue() const { return myTile->isConstant() ? myTile->constantValue() : myTile->rawData()[myIdx]; }
    void setValue(T v) { myTile->set(myLocal[0], myLocal[1], myLocal[2], v); }
    int x() const { return myTileOrigin[0] + myLocal[0]; }
    int y() const { return myTileOrigin[1] + myLocal[1]; }
    int z() const { return myTileOrigin[2] + myLocal[2]; }

private:
    UT_VoxelTile<T> *myTile = nullptr;
    int myLocal[3] = {0, 0, 0};
    int myTileOrigin[3] = {0, 0, 0};
    int myTileDim[3] = {0, 0, 0};
    int myIdx = 0;
    bool myDone = true;
};

// ---------------------------------------------------------------- probes
// The HDK probes cache an x-row of the tile under the index; reads and writes
// through them are semantically direct array accesses, which is what these do
// (with a same-tile fast path).
template <typename T, bool DoRead, bool DoWrite, bool TestForWrites>
class UT_VoxelProbe
{
public:
    void setArray(UT_VoxelArray<T> *a, int = 0, int = 0) { myArray = a; myTileNum = -1; }
    void setConstArray(const UT_VoxelArray<T> *a, int = 0, int = 0) { myArray = const_cast<UT_VoxelArray<T> *>(a); myTileNum = -1; }

    bool setIndex(int x, int y, int z)
    {
	myX = x; myY = y; myZ = z;
	if (x >= 0 && y >= 0 && z >= 0 && x < myArray->getXRes() && y < myArray->getYRes() && z < myArray->getZRes())
	{
	    int t = myArray->indexToLinearTile(x, y, z);
	    if (t != myTileNum) { myTileNum = t; myTile = myArray->getLinearTile(t); }
	    myInside = true;
	}
	else myInside = false;
	return true;
    }
    template <typename S>
    bool setIndex(const UT_VoxelArrayIterator<S> &vit) { return setIndex(vit.x(), vit.y(), vit.z()); }

    T getValue() const { return getValue(0); }
    T getValue(int dx) const
    {
	if (myInside)
	{
	    const int lx = (myX & GMG_TILEMASK) + dx;
	    if (lx >= 0 && lx < myTile->xres())
		return myTile->get(lx, myY & GMG_TILEMASK, myZ & GMG_TILEMASK);
	}
	return (*myArray)(myX + dx, myY, myZ);
    }
    void setValue(T v)
    {
	assert(myInside);
	myTile->set(myX & GMG_TILEMASK, myY & GMG_TILEMASK, myZ & GMG_TILEMASK, v);
    }

private:
    UT_VoxelArray<T> *myArray = nullptr;
    UT_VoxelTile<T> *myTile = nullptr;
    int myTileNum = -1;
    int myX = 0, myY = 0, myZ = 0;
    bool myInside = false;
};

template <typename T>
class UT_VoxelProbeCube
{
public:
    void setConstPlusArray(const UT_VoxelArray<T> *a) { myArray = a; myTileNum = -1; }
    void setConstCubeArray(const UT_VoxelArray<T> *a) { myArray = a; myTileNum = -1; }

    bool setIndexPlus(int x, int y, int z)
    {
	myX = x; myY = y; myZ = z;
	const int lx = x & GMG_TILEMASK, ly = y & GMG_TILEMASK, lz = z & GMG_TILEMASK;
	myFast = false;
	if (x >= 0 && y >= 0 && z >= 0 && x < myArray->getXRes() && y < myArray->getYRes() && z < myArray->getZRes())
	{
	    int t = myArray->indexToLinearTile(x, y, z);
	    if (t != myTileNum) { myTileNum = t; myTile = myArray->getLinearTile(t); }
	    myFast = lx > 0 && ly > 0 && lz > 0 && lx < myTile->xres() - 1 && ly < myTile->yres() - 1 && lz < myTile->zres() - 1;
	}
	return true;
    }
    template <typename S>
    bool setIndexPlus(const UT_VoxelArrayIterator<S> &vit) { return setIndexPlus(vit.x(), vit.y(), vit.z()); }
    bool setIndexCube(int x, int y, int z) { return setIndexPlus(x, y, z); }

    T getValue(int dx, int dy, int dz) const
    {
	if (myFast)
	    return myTile->get((myX & GMG_TILEMASK) + dx, (myY & GMG_TILEMASK) + dy, (myZ & GMG_TILEMASK) + dz);
	return (*myArray)(myX + dx, myY + dy, myZ + dz);
    }
    T getValue(const UT_Vector3I &o) const { return getValue(int(o[0]), int(o[1]), int(o[2])); }

private:
    const UT_VoxelArray<T> *myArray = nullptr;
    const UT_VoxelTile<T> *myTile = nullptr;
    int myTileNum = -1;
    int myX = 0, myY = 0, myZ = 0;
    bool myFast = false;
};

// ---------------------------------------------------------------- SIM_RawField / SIM_RawIndexField
// The slice HDK_Utilities.{h,cpp} uses (buildMaterialCellLabels, classifyValidFaces, ...): a voxel array of fpreal32 / exint
// with a resolution, tile access and -- SIM_RawField only -- a sample type (cell centres, or the faces of one axis), the box its
// cell lattice spans, indexToPos and a trilinear getValue(pos) clamped at the border.  At a sample position the interpolation
// weights are exactly 0 and 1, so getValue(indexToPos(cell)) of an ALIGNED field is that field's own value: the case the
// product's entry points document (gmg_build_material_labels, gmg_build_rhs).
using UT_VoxelArrayF = UT_VoxelArray<fpreal32>;
using UT_VoxelArrayI = UT_VoxelArray<exint>;
using UT_VoxelArrayIteratorF = UT_VoxelArrayIterator<fpreal32>;
using UT_VoxelArrayIteratorI = UT_VoxelArrayIterator<exint>;
using UT_VoxelTileIteratorF = UT_VoxelTileIterator<fpreal32>;
using UT_VoxelTileIteratorI = UT_VoxelTileIterator<exint>;

enum SIM_FieldSample { SIM_SAMPLE_CENTER = 0, SIM_SAMPLE_FACEX, SIM_SAMPLE_FACEY, SIM_SAMPLE_FACEZ };

class SIM_RawField
{
public:
    // (x, y, z) = resolution of the CELL lattice; a face-sampled field has one more entry along its axis.  orig / size: the box
    // the cell lattice spans (default: the unit-spaced lattice starting at 0).
    void init(SIM_FieldSample sample, const UT_Vector3 &orig, const UT_Vector3 &size, int x, int y, int z)
    {
	mySample = sample;
	myOrig = orig;
	mySize = size;
	myCells = UT_Vector3I(x, y, z);
	const int fa = faceAxis();
	myField.size(x + (fa == 0), y + (fa == 1), z + (fa == 2));
	myField.constant(0);
    }
    void init(int x, int y, int z) { init(SIM_SAMPLE_CENTER, UT_Vector3(0, 0, 0), UT_Vector3(fpreal32(x), fpreal32(y), fpreal32(z)), x, y, z); }
    void match(const SIM_RawField &o) { init(o.mySample, o.myOrig, o.mySize, int(o.myCells[0]), int(o.myCells[1]), int(o.myCells[2])); }
    void makeConstant(fpreal32 v) { myField.constant(v); }
    // HDK: "the resolution of the voxel grid that we are sampling" -- the CELL lattice, whatever the sample type (the reference asserts it equal
    // between a cell field and a face field, HDK_Utilities.h:89, :148); the array itself (field()) has one more entry along a face field's axis
    UT_Vector3I getVoxelRes() const { return myCells; }
    const UT_VoxelArrayF *field() const { return &myField; }
    UT_VoxelArrayF *fieldNC() { return &myField; }
    SIM_FieldSample getSample() const { return mySample; }
    const UT_Vector3 &getOrig() const { return myOrig; }
    const UT_Vector3 &getSize() const { return mySize; }
    UT_Vector3 getVoxelSize() const { return UT_Vector3(mySize[0] / fpreal32(myCells[0]), mySize[1] / fpreal32(myCells[1]), mySize[2] / fpreal32(myCells[2])); }
    bool isAligned(const SIM_RawField *o) const
    {
	return mySample == o->mySample && myCells == o->myCells && myOrig == o->myOrig && mySize == o->mySize;
    }
    // centre samples sit at (index + 0.5) dx, the samples of a face field at index dx along its own axis
    bool indexToPos(int x, int y, int z, UT_Vector3 &pos) const
    {
	const UT_Vector3 dx = getVoxelSize();
	const int idx[3] = {x, y, z};
	const int fa = faceAxis();
	for (int a = 0; a < 3; ++a) pos[a] = myOrig[a] + (fpreal32(idx[a]) + (a == fa ? 0.f : 0.5f)) * dx[a];
	return true;
    }
    // HDK: cut-cell face fractions of an SDF -- a geometric routine of the HDK that is not restated.  Only the diagnostic node's solid-sphere
    // option calls it (HDK_TestGeometricMultigrid.cpp:320), which the oracle never enables.
    void computeSDFWeightsFace(const SIM_RawField *, int, bool, fpreal = 0)
    {
	std::cerr << "hdk_shim: SIM_RawField::computeSDFWeightsFace is not available (useSolidSphere is off-path)" << std::endl;
	std::abort();
    }
    // trilinear, clamped at the border; at a sample position the weights are exactly 0 and 1
    fpreal32 getValue(const UT_Vector3 &pos) const
    {
	int i0[3];
	fpreal32 f[3];
	const UT_Vector3I r = myField.getVoxelRes();
	const UT_Vector3 dx = getVoxelSize();
	const int fa = faceAxis();
	for (int a = 0; a < 3; ++a)
	{
	    fpreal32 p = (pos[a] - myOrig[a]) / dx[a] - (a == fa ? 0.f : 0.5f);
	    p = p < 0 ? 0 : (p > fpreal32(r[a] - 1) ? fpreal32(r[a] - 1) : p);
	    i0[a] = int(p);
	    f[a] = p - fpreal32(i0[a]);
	}
	auto at = [&](int ddx, int ddy, int ddz) { return myField(i0[0] + ddx, i0[1] + ddy, i0[2] + ddz); };  // clamped read
	auto lerp = [](fpreal32 v0, fpreal32 v1, fpreal32 t) { return (1 - t) * v0 + t * v1; };
	return lerp(lerp(lerp(at(0, 0, 0), at(1, 0, 0), f[0]), lerp(at(0, 1, 0), at(1, 1, 0), f[0]), f[1]),
		    lerp(lerp(at(0, 0, 1), at(1, 0, 1), f[0]), lerp(at(0, 1, 1), at(1, 1, 1), f[0]), f[1]), f[2]);
    }

private:
    int faceAxis() const { return mySample == SIM_SAMPLE_CENTER ? -1 : int(mySample) - int(SIM_SAMPLE_FACEX); }
    UT_VoxelArrayF myField;
    SIM_FieldSample mySample = SIM_SAMPLE_CENTER;
    UT_Vector3 myOrig = UT_Vector3(0, 0, 0), mySize = UT_Vector3(1, 1, 1);
    UT_Vector3I myCells = UT_Vector3I(1, 1, 1);
};

class SIM_RawIndexField
{
public:
    void init(int x, int y, int z) { myField.size(x, y, z); myField.constant(0); }
    void match(const SIM_RawField &o) { const UT_Vector3I r = o.getVoxelRes(); init(int(r[0]), int(r[1]), int(r[2])); }
    void match(const SIM_RawIndexField &o) { const UT_Vector3I r = o.getVoxelRes(); init(int(r[0]), int(r[1]), int(r[2])); }
    void makeConstant(exint v) { myField.constant(v); }
    UT_Vector3I getVoxelRes() const { return myField.getVoxelRes(); }
    const UT_VoxelArrayI *field() const { return &myField; }
    UT_VoxelArrayI *fieldNC() { return &myField; }

private:
    UT_VoxelArrayI myField;
};

// ---------------------------------------------------------------- SIM::FieldUtils
namespace SIM
{
namespace FieldUtils
{
    inline fpreal32 getFieldValue(const SIM_RawField &field, const UT_Vector3I &cell) { return (*field.field())(cell); }
    inline exint getFieldValue(const SIM_RawIndexField &field, const UT_Vector3I &cell) { return (*field.field())(cell); }
    inline void setFieldValue(SIM_RawField &field, const UT_Vector3I &cell, const fpreal32 value) { field.fieldNC()->setValue(cell, value); }
    inline void setFieldValue(SIM_RawIndexField &field, const UT_Vector3I &cell, const exint value) { field.fieldNC()->setValue(cell, value); }

    // cell -> neighbouring cell along axis; direction 0 = backward, 1 = forward
    SYS_FORCE_INLINE UT_Vector3I cellToCellMap(const UT_Vector3I &cell, const int axis, const int direction)
    {
	UT_Vector3I adjacent(cell);
	if (direction == 0) --adjacent[axis];
	else ++adjacent[axis];
	return adjacent;
    }
    // cell -> its face along axis; backward face shares the cell index
    SYS_FORCE_INLINE UT_Vector3I cellToFaceMap(const UT_Vector3I &cell, const int axis, const int direction)
    {
	UT_Vector3I face(cell);
	if (direction == 1) ++face[axis];
	return face;
    }
    // face -> cell behind (direction 0) or in front (direction 1) of it
    SYS_FORCE_INLINE UT_Vector3I faceToCellMap(const UT_Vector3I &face, const int axis, const int direction)
    {
	UT_Vector3I cell(face);
	if (direction == 0) --cell[axis];
	return cell;
    }
    // Named by using-declarations in the reference (Ops.h:1466-1467) but never
    // called on the hot path; declared over plain voxel arrays so they resolve.
    template <typename T>
    inline T getFieldValue(const UT_VoxelArray<T> &field, const UT_Vector3I &cell) { return field(cell); }
    template <typename T>
    inline void setFieldValue(UT_VoxelArray<T> &field, const UT_Vector3I &cell, const T value) { field.setValue(cell, value); }

    template <typename Body>
    inline void forEachVoxelRange(const UT_Vector3I &start, const UT_Vector3I &end, const Body &body)
    {
	UT_Vector3I cell;
	for (cell[0] = start[0]; cell[0] < end[0]; ++cell[0])
	    for (cell[1] = start[1]; cell[1] < end[1]; ++cell[1])
		for (cell[2] = start[2]; cell[2] < end[2]; ++cell[2])
		    body(cell);
    }
} // namespace FieldUtils
} // namespace SIM

#endif
