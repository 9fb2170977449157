// ref_testnode_bridge.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the reference's own DIAGNOSTIC node, HDK_TestGeometricMultigrid.cpp (compiled unmodified from /root/reference/Source over
// oracle/shim), and hands its log back: the symmetry battery of the operators, smoothers, transfer pair, direct solve, one-level and
// full V-cycle (Test.cpp:1165-1875), the MGPCG test on the reference's own domains with its own delta right-hand side (Test.cpp:675-1163),
// the one-level V-cycle and smoother convergence tests.  The node builds its own inputs, so an entry point takes options only.
// A library of its own (oracle/_ref/libgmg_ref_testnode.so): both nodes define the DSO entry point initializeSIM.
#include <cstring>
#include <sstream>
#include <string>

#include "HDK_GeometricCGPoissonSolver.h"
#include "HDK_GeometricMultigridOperators.h"
#include "HDK_GeometricMultigridPoissonSolver.h"
#include "HDK_Utilities.h"
#include "hdk_node_shim.h"
// the node's constructor and solveGasSubclass are protected: read the declaration with every member public (see ref_bridge.cpp)
#define private public
#define protected public
#include "HDK_TestGeometricMultigrid.h"
#undef protected
#undef private

extern "C"
{
// options: "name=value;name=value;..." with the node's DOP parameter names (HDK_TestGeometricMultigrid.h:10-35), e.g.
// "gridSize=32;useComplexDomain=1;testSymmetry=1".  Returns 1 if the node reported success; the tail of its log goes to `log`.
int ref_testnode_run(const char *options, char *log, int logCap)
{
    HDK_TestGeometricMultigrid node(nullptr);
    std::string text(options ? options : "");
    size_t pos = 0;
    while (pos < text.size())
    {
	size_t end = text.find(';', pos);
	if (end == std::string::npos) end = text.size();
	const std::string item = text.substr(pos, end - pos);
	const size_t eq = item.find('=');
	if (eq != std::string::npos) node.options[item.substr(0, eq)] = std::atof(item.substr(eq + 1).c_str());
	pos = end + 1;
    }
    SIM_Object obj;
    SIM_Engine engine;
    std::ostringstream captured;
    std::streambuf *old = std::cout.rdbuf(captured.rdbuf());
    const bool ok = node.solveGasSubclass(engine, &obj, 0, 1. / 24.);
    std::cout.rdbuf(old);
    std::string out = captured.str();
    for (const std::string &e : obj.errors) out += "ERROR: " + e + "\n";
    if (log && logCap > 0)
    {
	const size_t n = std::min(out.size(), size_t(logCap - 1));
	std::memcpy(log, out.data() + (out.size() - n), n);
	log[n] = 0;
    }
    return ok ? 1 : 0;
}
void ref_testnode_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}
} // extern "C"
