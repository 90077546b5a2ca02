//go:build cuda

// Package conv -- cgo backend for dsp/conv on top of libalgodsp_cuda (include/algodsp_cuda.h).
//
// Drop this file into dsp/conv of github.com/cwbudde/algo-dsp and build with `-tags cuda`; the
// pure-Go files get `//go:build !cuda`.  Every exported signature below is the reference's
// (dsp/conv/conv.go, overlap_add.go, overlap_save.go, correlate.go, partitioned.go); the
// sentinel errors are the reference's own variables, so errors.Is keeps working.
//
// NOTE: this image has no Go toolchain, so this file has not been compiled here; it is kept
// deliberately thin (argument marshalling only -- every rule lives in the C library).
package conv

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -lalgodsp_cuda
#include <stdlib.h>
#include "algodsp_cuda.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"math"
	"runtime"
	"sync"
	"unsafe"
)

var (
	ErrEmptyInput           = errors.New("conv: empty input")
	ErrEmptyKernel          = errors.New("conv: empty kernel")
	ErrLengthMismatch       = errors.New("conv: buffer length mismatch")
	ErrInvalidBlockSize     = errors.New("conv: invalid block size")
	ErrInvalidBlockOrder    = errors.New("conv: invalid block order")
	ErrEmptyImpulseResponse = errors.New("conv: empty impulse response")
	ErrStageIndexOutOfRange = errors.New("conv: stage index out of range")
	ErrCUDA                 = errors.New("conv: CUDA error")
)

type Mode int

const (
	ModeFull Mode = iota
	ModeSame
	ModeValid
)

var (
	ctxOnce sync.Once
	ctx     *C.adsp_ctx
	ctxErr  error
)

func context() (*C.adsp_ctx, error) {
	ctxOnce.Do(func() {
		if st := C.adsp_ctx_create(0, &ctx); st != C.ADSP_OK {
			ctxErr = statusErr(st)
		}
	})
	return ctx, ctxErr
}

func statusErr(st C.adsp_status) error {
	switch st {
	case C.ADSP_OK:
		return nil
	case C.ADSP_ERR_EMPTY_INPUT:
		return ErrEmptyInput
	case C.ADSP_ERR_EMPTY_KERNEL:
		return ErrEmptyKernel
	case C.ADSP_ERR_LENGTH_MISMATCH:
		return fmt.Errorf("%w: %s", ErrLengthMismatch, lastError())
	case C.ADSP_ERR_INVALID_BLOCK_SIZE:
		return fmt.Errorf("%w: %s", ErrInvalidBlockSize, lastError())
	case C.ADSP_ERR_INVALID_BLOCK_ORDER:
		return fmt.Errorf("%w: %s", ErrInvalidBlockOrder, lastError())
	case C.ADSP_ERR_EMPTY_IR:
		return ErrEmptyImpulseResponse
	case C.ADSP_ERR_STAGE_INDEX:
		return fmt.Errorf("%w: %s", ErrStageIndexOutOfRange, lastError())
	default:
		return fmt.Errorf("%w: %s", ErrCUDA, lastError())
	}
}

func lastError() string {
	buf := make([]byte, 512)
	n := C.adsp_last_error((*C.char)(unsafe.Pointer(&buf[0])), C.size_t(len(buf)))
	if int(n) > len(buf)-1 {
		n = C.size_t(len(buf) - 1)
	}
	return string(buf[:n])
}

func ptr(s []float64) *C.double {
	if len(s) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&s[0]))
}

type binaryFn func(*C.adsp_ctx, *C.double, C.int64_t, *C.double, C.int64_t, *C.double) C.adsp_status

func binary(fn binaryFn, a, b []float64) ([]float64, error) {
	c, err := context()
	if err != nil {
		return nil, err
	}
	n := len(a) + len(b) - 1
	if n < 1 {
		n = 1
	}
	out := make([]float64, n)
	// host pointers are only used during the call (cgo pointer rule); the call is synchronous
	if st := fn(c, ptr(a), C.int64_t(len(a)), ptr(b), C.int64_t(len(b)), ptr(out)); st != C.ADSP_OK {
		return nil, statusErr(st)
	}
	return out[:len(a)+len(b)-1], nil
}

// Direct -- conv.go:76.
func Direct(a, b []float64) ([]float64, error) {
	return binary(func(c *C.adsp_ctx, a *C.double, n C.int64_t, b *C.double, m C.int64_t, o *C.double) C.adsp_status {
		return C.adsp_direct(c, a, n, b, m, o)
	}, a, b)
}

// Convolve -- conv.go:194 (auto-select lives in the C library: direct iff len(shorter) <= 64).
func Convolve(a, b []float64) ([]float64, error) {
	return binary(func(c *C.adsp_ctx, a *C.double, n C.int64_t, b *C.double, m C.int64_t, o *C.double) C.adsp_status {
		return C.adsp_convolve(c, a, n, b, m, o)
	}, a, b)
}

// ConvolveMode -- conv.go:219.
func ConvolveMode(a, b []float64, mode Mode) ([]float64, error) {
	full, err := Convolve(a, b)
	if err != nil {
		return nil, err
	}
	var s, l C.int64_t
	C.adsp_trim_mode(C.int64_t(len(a)), C.int64_t(len(b)), C.adsp_mode(mode), &s, &l)
	return full[s : s+l], nil
}

// Correlate -- correlate.go:16.
func Correlate(a, b []float64) ([]float64, error) {
	return binary(func(c *C.adsp_ctx, a *C.double, n C.int64_t, b *C.double, m C.int64_t, o *C.double) C.adsp_status {
		return C.adsp_correlate(c, a, n, b, m, o)
	}, a, b)
}

// AutoCorrelate -- correlate.go:57.
func AutoCorrelate(a []float64) ([]float64, error) { return Correlate(a, a) }

// FindPeak -- correlate.go:200.
func FindPeak(corr []float64) (index int, value float64) {
	c, err := context()
	if err != nil || len(corr) == 0 {
		return -1, 0
	}
	var idx C.int64_t
	var val C.double
	C.adsp_find_peak(c, ptr(corr), C.int64_t(len(corr)), &idx, &val)
	return int(idx), float64(val)
}

// LagFromIndex / IndexFromLag -- correlate.go:221-229.
func LagFromIndex(index, lenB int) int { return index - (lenB - 1) }
func IndexFromLag(lag, lenB int) int   { return lag + (lenB - 1) }

// OverlapSave -- overlap_save.go:32.  The kernel spectrum lives on the device inside the plan.
type OverlapSave struct{ plan *C.adsp_plan }

// NewOverlapSave -- overlap_save.go:53.
func NewOverlapSave(kernel []float64, fftSize int) (*OverlapSave, error) {
	c, err := context()
	if err != nil {
		return nil, err
	}
	os := &OverlapSave{}
	if st := C.adsp_overlap_save_create(c, unsafe.Pointer(ptr(kernel)), C.int64_t(len(kernel)), C.int64_t(fftSize), C.ADSP_F64, &os.plan); st != C.ADSP_OK {
		return nil, statusErr(st)
	}
	runtime.SetFinalizer(os, func(o *OverlapSave) { o.Close() })
	return os, nil
}

func (os *OverlapSave) Close() {
	if os.plan != nil {
		C.adsp_plan_destroy(os.plan)
		os.plan = nil
	}
}
func (os *OverlapSave) FFTSize() int   { return int(C.adsp_plan_fft_size(os.plan)) }
func (os *OverlapSave) StepSize() int  { return int(C.adsp_plan_step_size(os.plan)) }
func (os *OverlapSave) KernelLen() int { return int(C.adsp_plan_kernel_len(os.plan)) }
func (os *OverlapSave) Reset()         { C.adsp_plan_reset(os.plan) }

// Process -- overlap_save.go:126.
func (os *OverlapSave) Process(input []float64) ([]float64, error) {
	if len(input) == 0 {
		return nil, ErrEmptyInput
	}
	out := make([]float64, len(input)+os.KernelLen()-1)
	if st := C.adsp_plan_process(os.plan, unsafe.Pointer(ptr(input)), C.int64_t(len(input)), unsafe.Pointer(ptr(out)), C.int64_t(len(out))); st != C.ADSP_OK {
		return nil, statusErr(st)
	}
	return out, nil
}

// ProcessTo -- overlap_save.go:258.
func (os *OverlapSave) ProcessTo(output, input []float64) error {
	return statusErr(C.adsp_plan_process(os.plan, unsafe.Pointer(ptr(input)), C.int64_t(len(input)), unsafe.Pointer(ptr(output)), C.int64_t(len(output))))
}

// Deconvolution -- deconvolve.go:20-434.
type DeconvMethod int

const (
	DeconvNaive DeconvMethod = iota
	DeconvRegularized
	DeconvWiener
)

type DeconvOptions struct {
	Method         DeconvMethod
	Epsilon        float64
	NoiseVariance  float64
	SignalVariance float64
}

var ErrDivisionByZero = errors.New("conv: division by zero in deconvolution")

func DefaultDeconvOptions() DeconvOptions { return DeconvOptions{Method: DeconvRegularized, Epsilon: 1e-6} }

// Deconvolve -- deconvolve.go:72.
func Deconvolve(signal, kernel []float64, opts DeconvOptions) ([]float64, error) {
	if len(signal) == 0 {
		return nil, ErrEmptyInput
	}
	if len(kernel) == 0 {
		return nil, ErrEmptyKernel
	}
	c, err := context()
	if err != nil {
		return nil, err
	}
	out := make([]float64, int(C.adsp_deconv_out_len(C.int64_t(len(signal)), C.int64_t(len(kernel)))))
	st := C.adsp_deconvolve(c, ptr(signal), C.int64_t(len(signal)), ptr(kernel), C.int64_t(len(kernel)),
		C.int(opts.Method), C.double(opts.Epsilon), C.double(opts.NoiseVariance), C.double(opts.SignalVariance),
		ptr(out), C.int64_t(len(out)))
	if st == C.ADSP_ERR_DIVISION_BY_ZERO {
		return nil, fmt.Errorf("%w: %s", ErrDivisionByZero, lastError())
	}
	if st != C.ADSP_OK {
		return nil, statusErr(st)
	}
	return out, nil
}

// InverseFilter -- deconvolve.go:359.
func InverseFilter(kernel []float64, length int, epsilon float64) ([]float64, error) {
	if len(kernel) == 0 {
		return nil, ErrEmptyKernel
	}
	c, err := context()
	if err != nil {
		return nil, err
	}
	out := make([]float64, length)
	if length > 0 {
		if st := C.adsp_inverse_filter(c, ptr(kernel), C.int64_t(len(kernel)), C.int64_t(length), C.double(epsilon),
			ptr(out)); st != C.ADSP_OK {
			return nil, statusErr(st)
		}
	}
	return out, nil
}

// SNR -- deconvolve.go:417.
func SNR(original, recovered []float64) float64 {
	if len(original) != len(recovered) || len(original) == 0 {
		return math.Inf(-1)
	}
	return float64(C.adsp_snr(ptr(original), C.int64_t(len(original)), ptr(recovered), C.int64_t(len(recovered))))
}

// ConvolutionReverb backed by the device-resident frequency-domain delay line -- the reference type
// dsp/effects/reverb/convolution.go:17 keeps its API; `channels` > 1 is the GPU extension (rows of a block).
type ConvolutionReverb struct {
	plan     *C.adsp_plan
	channels int
}

// NewConvolutionReverb -- convolution.go:27 (maxBlockOrder 13).
func NewConvolutionReverb(kernel []float64, minBlockOrder, channels int) (*ConvolutionReverb, error) {
	if len(kernel) == 0 {
		return nil, errors.New("reverb: empty impulse response kernel")
	}
	c, err := context()
	if err != nil {
		return nil, err
	}
	var p *C.adsp_plan
	if st := C.adsp_partitioned_create_batch(c, unsafe.Pointer(ptr(kernel)), C.int64_t(len(kernel)), C.int(minBlockOrder), 13,
		C.int(channels), C.ADSP_F64, &p); st != C.ADSP_OK {
		return nil, statusErr(st)
	}
	r := &ConvolutionReverb{plan: p, channels: channels}
	runtime.SetFinalizer(r, func(r *ConvolutionReverb) { C.adsp_plan_destroy(r.plan) })
	return r, nil
}

// SetWetDry -- convolution.go:51.
func (r *ConvolutionReverb) SetWetDry(wet, dry float64) {
	C.adsp_partitioned_set_wet_dry(r.plan, C.double(wet), C.double(dry))
}

// ProcessInPlace -- convolution.go:60; block holds `channels` rows of len(block)/channels samples.
func (r *ConvolutionReverb) ProcessInPlace(block []float64) error {
	if len(block) == 0 {
		return nil
	}
	n := len(block) / r.channels
	return statusErr(C.adsp_partitioned_process_in_place_batch(r.plan, unsafe.Pointer(ptr(block)), C.int64_t(n), C.int64_t(n)))
}

func (r *ConvolutionReverb) Reset()       { C.adsp_plan_reset(r.plan) }                      // convolution.go:88
func (r *ConvolutionReverb) Latency() int { return int(C.adsp_partitioned_latency(r.plan)) } // convolution.go:98

// OverlapAdd, PartitionedConvolution, CorrelateFFT, CorrelateNormalized, ... follow the same
// pattern over adsp_overlap_add_create / adsp_partitioned_* / adsp_correlate_* (INTEGRATION.md).
