#define C2C_INST_VEC 4
#define C2C_INST_MASKED 0
#include "c2c_contract_inst.cuh"
