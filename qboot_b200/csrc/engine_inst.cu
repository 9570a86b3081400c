// qboot_b200 -- one translation unit per limb count: nvcc -DQB_L=<limbs> engine_inst.cu
#ifndef QB_L
#error "compile with -DQB_L=<number of 32-bit limbs>"
#endif
#include "engine.cuh"

#define QB_CAT2(a, b) a##b
#define QB_CAT(a, b) QB_CAT2(a, b)

extern "C" qb::EngineBase* QB_CAT(qb_new_engine_L, QB_L)(int device, int prec) { return qb::make_engine<QB_L>(device, prec); }
