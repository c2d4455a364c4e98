// Command golden dumps values computed by the UNTOUCHED reference packages (will-rowe/hulk v1.0.0) so that the
// CPU oracle of hulk-b200 (oracle/hulk_oracle.c, oracle/go_rand.c) can be pinned against the reference itself.
//
// No Go toolchain exists in the image hulk-b200 was built in, so this program has never been compiled there;
// it is the recipe for anyone who has Go (>= 1.12):
//
//	git clone https://github.com/will-rowe/hulk && cd hulk && git checkout v1.0.0
//	mkdir -p cmd/golden && cp <hulk-b200>/go/golden/main.go cmd/golden/main.go
//	go run ./cmd/golden -fastq testing/test-reads-small.fq.gz > <hulk-b200>/tests/golden/go_pins.json
//	cd <hulk-b200> && python -m pytest tests/test_go_pins.py -q
//
// (it has to live inside the hulk module because it imports the module's src/... packages.)
// tests/test_go_pins.py checks every field against the oracle when the file is present.
//
// What is dumped, and the reference code that computes it:
//   rng           leesper/go_rng GammaGenerator(1).Gamma(2,1) x 64, UniformGenerator(1).Float64Range(0,1) x 64
//                 -- the two streams HistoSketch.newCWS draws r, c, b from (src/histosketch/histosketch.go:103-116)
//   minimizers    per read of the fixture: NewMinimizerSketch(21, 9, seq).GetMinimizers(), sorted
//                 (src/minimizer/minimizer.go:59-204)
//   spectrum      KmerSpectrum of all those minimizers with 21^4 bins: used bins and md5 of the bin values as
//                 little-endian uint32 (src/kmerspectrum/kmerspectrum.go:67-81, dgryski/go-jump)
//   countmin      CountMinSketch.Add estimates for a fixed (bin, value) sequence, decay 1.0 and 0.02
//                 (src/countmin/countmin.go:103-147)
//   histosketch   Sketch / SketchWeights after feeding the spectrum's non-zero bins in ascending order once
//                 (k = 21, s = 50, decay 1.0 and 0.02): src/histosketch/histosketch.go:129-155 over the tables of newCWS
//   minhash       KMVsketch / KHFsketch (s = 50) after AddHash of every minimizer above, in read order -- the feed the
//                 reference's boss never makes (src/pipeline/boss.go:18-19) and hulk_b200_minhash_enable does
//                 (src/minhash/kmv.go:40-71,160-176, khf.go:35-45)
package main

import (
	"bufio"
	"compress/gzip"
	"crypto/md5"
	"encoding/binary"
	"encoding/json"
	"flag"
	"fmt"
	"math"
	"os"
	"sort"

	rng "github.com/leesper/go_rng"
	"github.com/will-rowe/hulk/src/countmin"
	"github.com/will-rowe/hulk/src/histosketch"
	"github.com/will-rowe/hulk/src/kmerspectrum"
	"github.com/will-rowe/hulk/src/minhash"
	"github.com/will-rowe/hulk/src/minimizer"
)

type cmsPin struct {
	Decay     float64   `json:"decay"`
	Bins      []uint64  `json:"bins"`
	Values    []float64 `json:"values"`
	Estimates []uint64  `json:"estimates_bits"` // math.Float64bits of every Add() result
}

type hskPin struct {
	Decay   float64  `json:"decay"`
	K       uint     `json:"k"`
	S       uint     `json:"s"`
	Mins    []uint64 `json:"mins"`
	Weights []uint64 `json:"weights_bits"`
}

type pins struct {
	Reference    string     `json:"reference"`
	GammaBits    []uint64   `json:"rng_gamma_2_1_bits"`
	UniformBits  []uint64   `json:"rng_uniform_0_1_bits"`
	K            uint       `json:"k"`
	W            uint       `json:"w"`
	NumBins      int32      `json:"num_bins"`
	Minimizers   [][]uint64 `json:"minimizers"`
	UsedBins     int        `json:"used_bins"`
	SpectrumMD5  string     `json:"spectrum_md5_u32le"`
	SpectrumBins []int32    `json:"spectrum_bins"`
	SpectrumFreq []float64  `json:"spectrum_freq"`
	CountMin     []cmsPin   `json:"countmin"`
	HistoSketch  []hskPin   `json:"histosketch"`
	KMV          []uint64   `json:"minhash_kmv"`
	KHF          []uint64   `json:"minhash_khf"`
}

func check(err error) {
	if err != nil {
		fmt.Fprintln(os.Stderr, "golden:", err)
		os.Exit(1)
	}
}

func readFastq(path string) [][]byte {
	fh, err := os.Open(path)
	check(err)
	defer fh.Close()
	gz, err := gzip.NewReader(fh)
	check(err)
	sc := bufio.NewScanner(gz)
	var seqs [][]byte
	for i := 0; sc.Scan(); i++ {
		if i%4 == 1 {
			seqs = append(seqs, append([]byte(nil), sc.Bytes()...))
		}
	}
	check(sc.Err())
	return seqs
}

func main() {
	fastq := flag.String("fastq", "testing/test-reads-small.fq.gz", "the reference's FASTQ fixture")
	flag.Parse()
	const k, w, s = uint(21), uint(9), uint(50)
	numBins := int32(k * k * k * k)
	out := pins{Reference: "will-rowe/hulk v1.0.0", K: k, W: w, NumBins: numBins}

	// the two generators of newCWS
	g := rng.NewGammaGenerator(histosketch.DISTRIBUTION_SEED)
	u := rng.NewUniformGenerator(histosketch.DISTRIBUTION_SEED)
	for i := 0; i < 64; i++ {
		out.GammaBits = append(out.GammaBits, math.Float64bits(g.Gamma(2, 1)))
		out.UniformBits = append(out.UniformBits, math.Float64bits(u.Float64Range(0, 1)))
	}

	// minimizers and the spectrum
	spectrum, err := kmerspectrum.NewKmerSpectrum(numBins)
	check(err)
	kmv, khf := minhash.NewKMVsketch(k, s), minhash.NewKHFsketch(k, s)
	for _, seq := range readFastq(*fastq) {
		ms, err := minimizer.NewMinimizerSketch(k, w, seq)
		check(err)
		var set []uint64
		for m := range ms.GetMinimizers() {
			set = append(set, m.(uint64))
			check(spectrum.AddHash(m.(uint64)))
			kmv.AddHash(m.(uint64))
			khf.AddHash(m.(uint64))
		}
		sort.Slice(set, func(a, b int) bool { return set[a] < set[b] })
		out.Minimizers = append(out.Minimizers, set)
	}
	out.UsedBins = spectrum.Cardinality()
	out.KMV, out.KHF = kmv.GetSketch(), khf.GetSketch()
	counts := make([]byte, 4*int(numBins))
	dump, err := spectrum.Dump()
	check(err)
	for bin := range dump {
		binary.LittleEndian.PutUint32(counts[4*int(bin.BinID):], uint32(bin.Frequency))
		out.SpectrumBins = append(out.SpectrumBins, bin.BinID)
		out.SpectrumFreq = append(out.SpectrumFreq, bin.Frequency)
	}
	out.SpectrumMD5 = fmt.Sprintf("%x", md5.Sum(counts))

	// count-min estimates for a fixed sequence (bins repeat, so estimates grow and, with decay, shrink)
	for _, decay := range []float64{1.0, 0.02} {
		cms := countmin.NewCountMinSketch(countmin.EPSILON, countmin.DELTA, decay)
		pin := cmsPin{Decay: decay}
		for i := 0; i < 400; i++ {
			bin := uint64((i*7919 + (i%5)*12345) % int(numBins))
			val := float64(1 + i%9)
			pin.Bins = append(pin.Bins, bin)
			pin.Values = append(pin.Values, val)
			pin.Estimates = append(pin.Estimates, math.Float64bits(cms.Add(bin, val)))
		}
		out.CountMin = append(out.CountMin, pin)
	}

	// the histosketch of the spectrum, one flush in ascending bin order (what theBoss.Flush does, boss.go:112-128)
	for _, decay := range []float64{1.0, 0.02} {
		hs, err := histosketch.NewHistoSketch(k, s, numBins, decay)
		check(err)
		for i, bin := range out.SpectrumBins {
			check(hs.AddElement(uint64(bin), out.SpectrumFreq[i]))
		}
		pin := hskPin{Decay: decay, K: k, S: s, Mins: hs.GetSketch()}
		for _, wgt := range hs.SketchWeights {
			pin.Weights = append(pin.Weights, math.Float64bits(wgt))
		}
		out.HistoSketch = append(out.HistoSketch, pin)
	}

	enc := json.NewEncoder(os.Stdout)
	check(enc.Encode(out))
}
