#define ADSP_REAL float
#include "fftconv_impl.cuh"
