// hdk_node_shim.h -- TEST INFRASTRUCTURE ONLY (oracle build).
//
// Stand-in for the slice of the Houdini 18 HDK that the reference's NODE source uses
// (HDK_GeometricFreeSurfacePressureSolver.{h,cpp}: a GAS_SubSolver with a DOP parameter description that pulls SIM fields off a
// SIM_Object, builds the solver's inputs, solves and writes pressure and velocity back), so that this file too compiles UNMODIFIED
// from /root/reference/Source into oracle/_ref/ and its functions -- buildMGDomainLabels, buildMGBoundaryWeights, buildRHS,
// applyOldPressure, applySolutionToPressure, applyPressureGradient, and the whole solveGasSubclass -- can be run on arrays.
//
// What is functional: SIM_Object is a bag of named fields, GAS_SubSolver looks fields and options up by name, SIM_ScalarField /
// SIM_VectorField own SIM_RawFields (hdk_shim.h) on one lattice.  What is inert: the parameter description (PRM_*,
// SIM_DopDescription), the data factory, the performance monitor, error reporting (collected as strings).
// Nothing in the shipped product includes this file.
#ifndef GMG_ORACLE_HDK_NODE_SHIM_H
#define GMG_ORACLE_HDK_NODE_SHIM_H

#include <map>
#include <random>
#include <string>

#include "hdk_shim.h"

#define GAS_API
using SIM_Time = fpreal64;

// ---------------------------------------------------------------- field containers
class SIM_ScalarField
{
public:
    SIM_RawField *getField() { return &myField; }
    const SIM_RawField *getField() const { return &myField; }
    void matchField(const SIM_ScalarField *o)
    {
	if (!(myField.getVoxelRes() == o->myField.getVoxelRes())) myField.match(o->myField);
    }
    void pubHandleModification() {}

private:
    SIM_RawField myField;
};

class SIM_VectorField
{
public:
    // cells: the resolution of the cell lattice; face sampled: component a has one more entry along a
    void initFaces(int x, int y, int z)
    {
	myCells = UT_Vector3I(x, y, z);
	for (int a = 0; a < 3; ++a) myFields[a].init(SIM_FieldSample(SIM_SAMPLE_FACEX + a), UT_Vector3(0, 0, 0), UT_Vector3(fpreal32(x), fpreal32(y), fpreal32(z)), x, y, z);
    }
    SIM_RawField *getField(int axis) { return &myFields[axis]; }
    const SIM_RawField *getField(int axis) const { return &myFields[axis]; }
    bool isFaceSampled() const { return myFields[0].getSample() == SIM_SAMPLE_FACEX && myFields[1].getSample() == SIM_SAMPLE_FACEY && myFields[2].getSample() == SIM_SAMPLE_FACEZ; }
    bool isAligned(const SIM_VectorField *o) const
    {
	for (int a = 0; a < 3; ++a)
	    if (!myFields[a].isAligned(&o->myFields[a])) return false;
	return true;
    }
    UT_Vector3I getTotalVoxelRes() const { return myCells; }
    UT_Vector3 getOrig() const { return myFields[0].getOrig(); }
    UT_Vector3 getSize() const { return myFields[0].getSize(); }
    UT_Vector3 getVoxelSize() const { return myFields[0].getVoxelSize(); }
    void pubHandleModification() {}

private:
    SIM_RawField myFields[3];
    UT_Vector3I myCells;
};

// ---------------------------------------------------------------- SIM_Object / engine / factory
class SIM_Engine {};
class SIM_DataFactory {};
class SIM_Object
{
public:
    std::map<std::string, SIM_ScalarField *> scalarFields;
    std::map<std::string, SIM_VectorField *> vectorFields;
    std::vector<std::string> errors;
};

enum { SIM_MESSAGE = 0 };
enum UT_ErrorSeverity { UT_ERROR_NONE = 0, UT_ERROR_MESSAGE, UT_ERROR_PROMPT, UT_ERROR_WARNING, UT_ERROR_ABORT, UT_ERROR_FATAL };

// ---------------------------------------------------------------- parameter description (inert)
enum PRM_Type { PRM_STRING, PRM_TOGGLE, PRM_FLT, PRM_INT, PRM_SEPARATOR };
class PRM_Conditional
{
public:
    PRM_Conditional(const char * = nullptr) {}
};
class PRM_Name
{
public:
    PRM_Name(const char * = nullptr, const char * = nullptr) {}
};
class PRM_Default
{
public:
    PRM_Default(fpreal = 0, const char * = nullptr) {}
};
static PRM_Default PRMoneDefaults[1] = {PRM_Default(1)};
static PRM_Default PRMzeroDefaults[1] = {PRM_Default(0)};
class PRM_Template
{
public:
    PRM_Template() {}
    // (type, vector size, name, defaults, choice list, range, callback, spare data, parameter group, help text, conditional)
    PRM_Template(PRM_Type, int, PRM_Name *, PRM_Default * = nullptr, void * = nullptr, void * = nullptr, void * = nullptr, void * = nullptr, int = 1,
		 const char * = nullptr, PRM_Conditional * = nullptr)
    {
    }
};
class SIM_DopDescription
{
public:
    SIM_DopDescription(bool, const char *, const char *, const char *, const char *, const PRM_Template *) {}
};

#define GAS_NAME_SURFACE "surface"
#define GAS_NAME_VELOCITY "velocity"
#define GAS_NAME_COLLISION "collision"
#define GAS_NAME_COLLISIONVELOCITY "collisionvel"
#define GAS_NAME_PRESSURE "pressure"
#define GAS_NAME_DENSITY "density"
#define SIM_NAME_TOLERANCE "tolerance"

// ---------------------------------------------------------------- GAS_SubSolver
class GAS_SubSolver
{
public:
    explicit GAS_SubSolver(const SIM_DataFactory *) {}
    virtual ~GAS_SubSolver() {}

    // options of the node (the DOP parameters), set by the test harness
    std::map<std::string, fpreal64> options;
    fpreal64 option(const char *name) const
    {
	auto it = options.find(name);
	return it == options.end() ? 0 : it->second;
    }

protected:
    SIM_ScalarField *getScalarField(SIM_Object *obj, const char *name, bool = false)
    {
	auto it = obj->scalarFields.find(name);
	return it == obj->scalarFields.end() ? nullptr : it->second;
    }
    const SIM_ScalarField *getConstScalarField(SIM_Object *obj, const char *name) { return getScalarField(obj, name); }
    SIM_VectorField *getVectorField(SIM_Object *obj, const char *name, bool = false)
    {
	auto it = obj->vectorFields.find(name);
	return it == obj->vectorFields.end() ? nullptr : it->second;
    }
    const SIM_VectorField *getConstVectorField(SIM_Object *obj, const char *name) { return getVectorField(obj, name); }
    void addError(SIM_Object *obj, int, const char *msg, UT_ErrorSeverity) { if (obj) obj->errors.push_back(msg); }
    static void setGasDescription(SIM_DopDescription &) {}
    virtual bool solveGasSubclass(SIM_Engine &, SIM_Object *, SIM_Time, SIM_Time) = 0;
};

#define GET_DATA_FUNC_F(DataName, FuncName) fpreal get##FuncName() const { return fpreal(option(DataName)); }
#define GET_DATA_FUNC_I(DataName, FuncName) int get##FuncName() const { return int(option(DataName)); }
#define GET_DATA_FUNC_B(DataName, FuncName) bool get##FuncName() const { return option(DataName) != 0; }
#define DECLARE_STANDARD_GETCASTTOTYPE()
#define DECLARE_DATAFACTORY(DataClass, SuperClass, Description, DopParms) \
public:                                                                   \
    typedef SuperClass BaseClass;                                         \
    static const char *classname() { return #DataClass; }                 \
private:
#define IMPLEMENT_DATAFACTORY(DataClass) (void)0

class UT_PerfMonAutoSolveEvent
{
public:
    template <typename Solver>
    UT_PerfMonAutoSolveEvent(const Solver *, const char *) {}
};

#endif
