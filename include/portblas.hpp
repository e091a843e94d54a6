// portblas.hpp -- header-only entry point (reference src/portblas.hpp:25-63).  In this build the
// "library" and "header-only" modes coincide: the templates are thin and the device code lives
// in libpbx_gemm.so, so this simply includes portblas.h.
#pragma once
#include "portblas.h"
