// gemb200_step_tu.cu — one translation unit per (motor family, real): compiled 10 times by build.py with
//   -DGEMB200_TU_FAM=<0..4> -DGEMB200_TU_REAL=<float|double>
#ifndef GEMB200_TU_FAM
#error "compile with -DGEMB200_TU_FAM=<family> -DGEMB200_TU_REAL=<float|double>"
#endif
#include "gemb200_launch.cuh"
