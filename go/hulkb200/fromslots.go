// UNCOMPILED: add to src/histosketch/ of will-rowe/hulk v1.0.0 (see INTEGRATION.md).
package histosketch

// FromSlots wraps slots computed outside this package (the GPU library owns the CWS/count-min state).
func FromSlots(k uint, dims int32, drift bool, mins []uint, weights []float64) *HistoSketch {
	return &HistoSketch{algorithm: "histosketch", KmerSize: k, SketchSize: uint(len(mins)), Dimensions: dims,
		ApplyConceptDrift: drift, Sketch: mins, SketchWeights: weights}
}
