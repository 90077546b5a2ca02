#define ADSP_REAL double
#include "fftconv_impl.cuh"
