// Library information entry points of the C ABI.
#include "common.cuh"

extern "C" {

int cb200_version(int* major, int* minor, int* sm_arch) {
  if (major) *major = 0;
  if (minor) *minor = 1;
  if (sm_arch) *sm_arch = 100;
  return CB200_OK;
}

const char* cb200_error_string(int code) {
  if (code == CB200_OK) return "ok";
  if (code == CB200_EINVAL) return "invalid argument";
  if (code == CB200_EUNSUPPORTED) return "unsupported dtype / dimensionality";
  if (code == CB200_ENOFIT) return "the fit subset is empty";
  if (code == CB200_ENOCENTRE) return "no point was within the bandwidth of any seed";
  if (code == CB200_ENOCONVERGE) return "centre suppression did not reach its fix-point";
  if (code == CB200_ENOSPACE) return "workspace too small for the counts found on the device (see info->workspace_needed)";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}

}  // extern "C"
