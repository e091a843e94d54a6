// interface/gemm_launcher.hpp -- the reference keeps the definition of Gemm_Launcher::_select_gemm here
// (src/interface/gemm_launcher.hpp:39-64); in this build the declaration carries it.
#pragma once
#include "interface/gemm_launcher.h"
