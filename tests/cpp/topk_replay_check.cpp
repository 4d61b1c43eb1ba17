// host build of pyg_lib_b200/csrc/topk_replay.h for tests/test_weighted_oracle.py (ctypes)
#include "../../pyg_lib_b200/csrc/topk_replay.h"
extern "C" void topk_replay(float* K, uint32_t* I, int n, int k) { w_topk_replay(WPairs{K, I}, n, k); }
