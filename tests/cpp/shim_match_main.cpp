// Test driver: the drop-in matcher classes (shim/ORBmatcher.h, shim/LSDmatcher.h, shim/FrustumGPU.h -> C ABI -> CUDA) behind the SAME
// driver and the same stand-in Frame / MapPoint / MapLine types the executed reference uses (oracle/ref_match_main.cpp): fed the same
// in.bin, the two binaries must write the same assignments (tests/test_shim_cpp.py).
#define REF_MATCH_USE_SHIM
#include "../../oracle/ref_match_main.cpp"
