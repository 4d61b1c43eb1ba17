// Shared bits of the torch registration layer (libpyg.so): the only place torch types appear.
// Everything below this layer is the C ABI of include/pyg_b200.h.
#pragma once
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include "pyg_b200.h"

#define PYGB_TORCH_CALL(expr)                                                        \
  do {                                                                               \
    const int _rc = (expr);                                                          \
    TORCH_CHECK(_rc == PYGB200_OK, "pyg_lib_b200: ", pygb200_last_error(), " [", #expr, "]"); \
  } while (0)
