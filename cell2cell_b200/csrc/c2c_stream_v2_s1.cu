#define C2C_INST_VEC 2
#define C2C_INST_NS 1
#include "c2c_stream_inst.cuh"
