// UNCOMPILED: no Go toolchain exists in the authoring image (see INTEGRATION.md).
// Drop into src/pipeline/ of will-rowe/hulk v1.0.0 and build with `-tags b200`.
// +build b200

package pipeline

/*
#cgo CFLAGS:  -I${SRCDIR}/../../third_party/hulk_b200/include
#cgo LDFLAGS: -L${SRCDIR}/../../third_party/hulk_b200/lib -lhulk_b200 -Wl,-rpath,$ORIGIN
#include <stdlib.h>
#include "hulk_b200.h"
*/
import "C"

import (
	"fmt"
	"log"
	"os"
	"runtime"
	"strconv"
	"unsafe"

	"github.com/will-rowe/hulk/src/helpers"
	"github.com/will-rowe/hulk/src/histosketch"
	"github.com/will-rowe/hulk/src/minhash"
	"github.com/will-rowe/hulk/src/seqio"
	"github.com/will-rowe/hulk/src/sketchio"
)

// GPUSketcher replaces SeqMinimizer + Sketcher (src/pipeline/sketch.go:163-301).
type GPUSketcher struct {
	info  *Info
	input chan *seqio.FASTQread
}

func NewGPUSketcher(info *Info) *GPUSketcher      { return &GPUSketcher{info: info} }
func (proc *GPUSketcher) Connect(p *FastqHandler) { proc.input = p.output }

const batchBytes = 64 << 20 // reads are handed to the GPU in ~64 MB batches

func (proc *GPUSketcher) Run() {
	runtime.LockOSThread() // a context is single-caller (include/hulk_b200.h)
	defer runtime.UnlockOSThread()
	s := proc.info.Sketch
	// One handle for 1..16 GPUs of this host (a group of one is the plain context): every interval's reads are split
	// over the GPUs, the spectra are summed over NVLink inside the flush, the sketch slots are sharded.
	// HULK_B200_F_PACK_INPUT: each batch crosses PCIe as 2 bits per base, packed by the library on this host's cores
	// inside push_reads (the Go slices below are free for reuse as soon as the call returns).
	ngpus := 1
	if v, err := strconv.Atoi(os.Getenv("HULK_B200_GPUS")); err == nil && v > 0 {
		ngpus = v
	}
	var g *C.hulk_b200_group
	p := C.hulk_b200_params{k: C.uint32_t(s.KmerSize), w: C.uint32_t(s.WindowSize),
		sketch_size: C.uint32_t(s.SketchSize), num_bins: C.int32_t(s.SpectrumSize),
		decay_ratio: C.double(s.DecayRatio), flags: C.HULK_B200_F_PACK_INPUT}
	check := func(rc C.int) {
		if rc != 0 {
			helpers.ErrorCheck(fmt.Errorf("%s", C.GoString(C.hulk_b200_group_last_error(g))))
		}
	}
	check(C.hulk_b200_group_create(&p, nil, C.uint32_t(ngpus), &g))
	defer C.hulk_b200_group_destroy(g)
	check(C.hulk_b200_group_generate_cws_tables(g, 1)) // newCWS (histosketch.go:95-126), drawn on all cores while reads are counted
	// --kmv / --khf: the reference's boss never feeds the two MinHash sketches (boss.go:18-19); with
	// HULK_B200_FEED_MINHASH=1 the library feeds them from the same minimizer stream (include/hulk_b200.h)
	feedMinhash := os.Getenv("HULK_B200_FEED_MINHASH") == "1" && (s.KMV || s.KHF)
	if feedMinhash {
		b2i := map[bool]C.int{false: 0, true: 1}
		check(C.hulk_b200_group_minhash_enable(g, b2i[s.KMV], b2i[s.KHF]))
	}

	bases := make([]byte, 0, batchBytes+1<<20)
	offsets := []C.uint64_t{0}
	push := func() {
		if len(offsets) > 1 {
			check(C.hulk_b200_group_push_reads(g, (*C.uint8_t)(unsafe.Pointer(&bases[0])),
				&offsets[0], C.uint64_t(len(offsets)-1)))
			bases, offsets = bases[:0], offsets[:1]
		}
	}
	log.Printf("finding minimizers...")
	seqCount, interval, sketchingInterval := uint(0), s.Interval, 0
	for sequence := range proc.input { // sketch.go:197
		bases = append(bases, sequence.Seq...)
		offsets = append(offsets, C.uint64_t(len(bases)))
		seqCount++
		if seqCount%100000 == 0 {
			log.Printf("\tprocessed %d sequences", seqCount)
		}
		if interval != 0 && seqCount%interval == 0 { // sketch.go:211-215
			push()
			sketchingInterval++
			log.Printf("\treached interval %d -> histosketching", sketchingInterval)
			check(C.hulk_b200_group_flush(g))
		} else if len(bases) >= batchBytes {
			push()
		}
	}
	log.Printf("generating final histosketch of k-mer spectra...")
	push()
	check(C.hulk_b200_group_flush(g)) // sketch.go:221
	if seqCount == 0 {
		helpers.ErrorCheck(fmt.Errorf("no sequences received")) // sketch.go:237-239
	}
	mins := make([]C.uint64_t, s.SketchSize)
	weights := make([]C.double, s.SketchSize)
	check(C.hulk_b200_group_finish(g, &mins[0], &weights[0]))
	var st C.hulk_b200_stats
	check(C.hulk_b200_group_get_stats(g, &st))
	log.Printf("\tprocessed %d sequences in total\n", seqCount)
	log.Printf("\tmean sequence length: %d\n", uint(float64(st.n_bases)/float64(seqCount)))
	log.Printf("\tfound %d minimizers\n", uint64(st.n_minimizers))
	log.Printf("\thistosketching across %d bins\n", s.SpectrumSize)

	// hand the result to the reference's own output code (sketchio.go:56-97); `algorithm`, `cwsSamples`
	// and `cmSketch` are unexported, so the histosketch package gains the 12-line constructor below
	sk := make([]uint, s.SketchSize)
	wt := make([]float64, s.SketchSize)
	for i := range mins {
		sk[i], wt[i] = uint(mins[i]), float64(weights[i])
	}
	hs := histosketch.FromSlots(s.KmerSize, s.SpectrumSize, s.DecayRatio != 1.0, sk, wt)
	hulkData := sketchio.NewHULKdata()
	helpers.ErrorCheck(hulkData.Add(hs))
	if feedMinhash && s.KMV { // sketch.go:227-230: KMVsketch.AddHash keeps what it is given when it fits (kmv.go:57-59)
		vals := make([]C.uint64_t, s.SketchSize)
		var n C.uint32_t
		check(C.hulk_b200_group_get_kmv(g, &vals[0], &n))
		kmv := minhash.NewKMVsketch(s.KmerSize, s.SketchSize)
		for _, v := range vals[:n] {
			kmv.AddHash(uint64(v))
		}
		helpers.ErrorCheck(hulkData.Add(kmv))
	}
	if feedMinhash && s.KHF { // sketch.go:231-234
		vals := make([]C.uint64_t, s.SketchSize)
		check(C.hulk_b200_group_get_khf(g, &vals[0]))
		khf := minhash.NewKHFsketch(s.KmerSize, s.SketchSize)
		for i, v := range vals {
			khf.Sketch[i] = uint64(v)
		}
		helpers.ErrorCheck(hulkData.Add(khf))
	}
	hulkData.FileName, hulkData.Banner = s.FileName, s.BannerLabel
	hulkData.WriteJSON(s.OutFile + ".json")
	log.Printf("\twritten sketch to disk: %v\n", s.OutFile+".json")
}
