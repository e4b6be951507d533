// Hand-written stand-in for the AMReX_buildInfo.cpp that AMReX's makebuildinfo_C.py would generate (Exec/Make.PeleC:186-192):
// PeleC only prints these strings into job_info (Source/IO.cpp:490-520).  Test infrastructure (oracle/ref_pelec).
#include <AMReX_buildInfo.H>
namespace amrex {
const char* buildInfoGetBuildDate() { return "n/a"; }
const char* buildInfoGetBuildDir() { return "oracle/ref_pelec"; }
const char* buildInfoGetBuildMachine() { return "n/a"; }
const char* buildInfoGetAMReXDir() { return "Submodules/PelePhysics/Submodules/amrex"; }
const char* buildInfoGetComp() { return "gnu"; }
const char* buildInfoGetCompVersion() { return __VERSION__; }
const char* buildInfoGetCompilerLoc(int) { return ""; }
const char* buildInfoGetCompilerVer(int) { return ""; }
const char* buildInfoGetCompilerFlags(int) { return ""; }
const char* buildInfoGetFcomp() { return ""; }
const char* buildInfoGetFcompVersion() { return ""; }
const char* buildInfoGetCXXName() { return "g++"; }
const char* buildInfoGetFName() { return ""; }
const char* buildInfoGetCXXFlags() { return ""; }
const char* buildInfoGetFFlags() { return ""; }
const char* buildInfoGetLinkFlags() { return ""; }
const char* buildInfoGetLibraries() { return ""; }
const char* buildInfoGetAux(int) { return ""; }
int buildInfoGetNumModules() { return 0; }
const char* buildInfoGetModuleName(int) { return ""; }
const char* buildInfoGetModuleVal(int) { return ""; }
const char* buildInfoGetDif(int) { return ""; }
const char* buildInfoGetGitHash(int) { return ""; }
const char* buildInfoGetGitBranch(int) { return ""; }
const char* buildInfoGetLinkerInfo(int) { return ""; }
const char* buildInfoGetBuildGitName() { return ""; }
const char* buildInfoGetBuildGitHash() { return ""; }
}
