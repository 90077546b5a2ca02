//go:build cuda

// cgo backends for the packages next to dsp/conv (SURVEY 8f #2-#4) on top of libalgodsp_cuda (include/algodsp_cuda.h):
// measure/sweep LogSweep, measure/ir Analyzer (Schroeder integral, impulse start), dsp/filter/fir Filter and
// dsp/resample Resampler.  Each block below belongs into the package named in its comment (they are collected in one
// file here because this image has no Go toolchain: the file is uncompiled and deliberately thin -- argument marshalling
// only, every rule lives in the C library).  Signatures are the reference's.
package conv

/*
#cgo LDFLAGS: -lalgodsp_cuda
#include "algodsp_cuda.h"
*/
import "C"

import (
	"errors"
	"runtime"
	"unsafe"
)

// RegisterBuffer pins a long-lived slice in place so that host-pointer calls DMA from / to it directly.  The caller keeps
// the slice alive and unmoved (runtime.Pinner) until UnregisterBuffer.
func RegisterBuffer(buf []float64, pin *runtime.Pinner) error {
	if len(buf) == 0 {
		return nil
	}
	pin.Pin(&buf[0])
	if st := C.adsp_host_register(unsafe.Pointer(&buf[0]), C.size_t(len(buf)*8)); st != C.ADSP_OK {
		return statusErr(st)
	}
	return nil
}

func UnregisterBuffer(buf []float64) error {
	if len(buf) == 0 {
		return nil
	}
	if st := C.adsp_host_unregister(unsafe.Pointer(&buf[0])); st != C.ADSP_OK {
		return statusErr(st)
	}
	return nil
}

// ---------------------------------------------------------------- package sweep (measure/sweep/sweep.go)
var ErrEmptyResponse = errors.New("sweep: response signal is empty")

type LogSweep struct {
	StartFreq, EndFreq, Duration, SampleRate float64
}

func (s *LogSweep) samples() int { return int(C.adsp_logsweep_samples(C.double(s.Duration), C.double(s.SampleRate))) }

// Generate -- sweep.go:73
func (s *LogSweep) Generate() ([]float64, error) {
	out := make([]float64, max(s.samples(), 1))
	if st := C.adsp_logsweep_generate_host((*C.double)(&out[0]), C.double(s.StartFreq), C.double(s.EndFreq), C.double(s.Duration), C.double(s.SampleRate)); st != C.ADSP_OK {
		return nil, statusErr(st)
	}
	return out[:s.samples()], nil
}

// InverseFilter -- sweep.go:104
func (s *LogSweep) InverseFilter() ([]float64, error) {
	out := make([]float64, max(s.samples(), 1))
	if st := C.adsp_logsweep_inverse_filter_host((*C.double)(&out[0]), C.double(s.StartFreq), C.double(s.EndFreq), C.double(s.Duration), C.double(s.SampleRate)); st != C.ADSP_OK {
		return nil, statusErr(st)
	}
	return out[:s.samples()], nil
}

// Deconvolve -- sweep.go:164
func (s *LogSweep) Deconvolve(response []float64) ([]float64, error) {
	if len(response) == 0 {
		return nil, ErrEmptyResponse
	}
	c, err := context()
	if err != nil {
		return nil, err
	}
	out := make([]float64, len(response)+s.samples()-1)
	st := C.adsp_logsweep_deconvolve(c, (*C.double)(&response[0]), C.int64_t(len(response)), C.double(s.StartFreq), C.double(s.EndFreq),
		C.double(s.Duration), C.double(s.SampleRate), (*C.double)(&out[0]), C.int64_t(len(out)))
	if st != C.ADSP_OK {
		return nil, statusErr(st)
	}
	return out, nil
}

// ---------------------------------------------------------------- package ir (measure/ir/ir.go)
var ErrEmptyIR = errors.New("ir: impulse response is empty")

type Analyzer struct{ SampleRate float64 }

// SchroederIntegral -- ir.go:94
func (a *Analyzer) SchroederIntegral(ir []float64) ([]float64, error) {
	if len(ir) == 0 {
		return nil, ErrEmptyIR
	}
	c, err := context()
	if err != nil {
		return nil, err
	}
	out := make([]float64, len(ir))
	if st := C.adsp_ir_schroeder(c, (*C.double)(&ir[0]), C.int64_t(len(ir)), (*C.double)(&out[0])); st != C.ADSP_OK {
		return nil, statusErr(st)
	}
	return out, nil
}

// FindImpulseStart -- ir.go:381 (threshold -20 dB re peak)
func (a *Analyzer) FindImpulseStart(ir []float64) (int, error) {
	if len(ir) == 0 {
		return 0, ErrEmptyIR
	}
	c, err := context()
	if err != nil {
		return 0, err
	}
	var idx C.int64_t
	if st := C.adsp_ir_find_impulse_start(c, (*C.double)(&ir[0]), C.int64_t(len(ir)), 0.1, &idx); st != C.ADSP_OK {
		return 0, statusErr(st)
	}
	return int(idx), nil
}

// ---------------------------------------------------------------- package fir (dsp/filter/fir/filter.go)
type Filter struct{ h *C.adsp_fir }

// New -- filter.go:18
func New(coeffs []float64) *Filter {
	c, err := context()
	if err != nil {
		return nil
	}
	f := &Filter{}
	var p *C.double
	if len(coeffs) > 0 {
		p = (*C.double)(&coeffs[0])
	}
	if st := C.adsp_fir_create(c, p, C.int64_t(len(coeffs)), 1, &f.h); st != C.ADSP_OK {
		return nil
	}
	runtime.SetFinalizer(f, func(f *Filter) { C.adsp_fir_destroy(f.h) })
	return f
}

// ProcessBlock -- filter.go:61 (in place, state carried across calls)
func (f *Filter) ProcessBlock(buf []float64) {
	if len(buf) > 0 {
		C.adsp_fir_process_block(f.h, (*C.double)(&buf[0]), C.int64_t(len(buf)), C.int64_t(len(buf)))
	}
}
func (f *Filter) Order() int { return int(C.adsp_fir_order(f.h)) }
func (f *Filter) Reset()     { C.adsp_fir_reset(f.h) }

// ---------------------------------------------------------------- package resample (dsp/resample/resample.go)
var ErrInvalidRatio = errors.New("resample: invalid ratio")

type Quality int

const (
	QualityFast Quality = iota
	QualityBalanced
	QualityBest
)

type Resampler struct{ h *C.adsp_resampler }

// NewRational -- resample.go:153 (quality option only; the other options map onto the trailing arguments)
func NewRational(up, down int, q Quality) (*Resampler, error) {
	if up <= 0 || down <= 0 {
		return nil, ErrInvalidRatio
	}
	c, err := context()
	if err != nil {
		return nil, err
	}
	r := &Resampler{}
	if st := C.adsp_resampler_create(c, C.int(up), C.int(down), C.int(q), 0, 0, 0, 1, &r.h); st != C.ADSP_OK {
		return nil, statusErr(st)
	}
	runtime.SetFinalizer(r, func(r *Resampler) { C.adsp_resampler_destroy(r.h) })
	return r, nil
}

// Process -- resample.go:249
func (r *Resampler) Process(input []float64) []float64 {
	if len(input) == 0 {
		return nil
	}
	n := int(C.adsp_resampler_predict_output_len(r.h, C.int64_t(len(input))))
	out := make([]float64, max(n, 1))
	var got C.int64_t
	C.adsp_resampler_process(r.h, (*C.double)(&input[0]), C.int64_t(len(input)), C.int64_t(len(input)), (*C.double)(&out[0]), C.int64_t(len(out)),
		C.int64_t(len(out)), &got)
	return out[:int(got)]
}
func (r *Resampler) PredictOutputLen(n int) int { return int(C.adsp_resampler_predict_output_len(r.h, C.int64_t(n))) }
func (r *Resampler) Ratio() (int, int) {
	var u, d C.int
	C.adsp_resampler_ratio(r.h, &u, &d)
	return int(u), int(d)
}
func (r *Resampler) Reset() { C.adsp_resampler_reset(r.h) }
