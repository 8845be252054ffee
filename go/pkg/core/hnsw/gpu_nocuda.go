//go:build !cuda

package hnsw

import (
	"errors"

	"github.com/RoaringBitmap/roaring"
	"github.com/sanonone/kektordb/pkg/core/types"
)

// Without the `cuda` build tag the hooks compile to nothing: SearchWithScores runs the CPU path of
// hnsw_index.go unchanged (same pattern as distance_go.go vs distance_rust.go).

type GPUOptions struct {
	Device         int
	MaxBatch       uint32
	MaxWaitMicros  uint32
	MaxPendingRows uint32
	MaxLagMillis   uint32
	CapacitySlack  uint32
}

func (h *Index) AttachGPU(GPUOptions) error {
	return errors.New("kektordb was built without the cuda tag")
}
func (h *Index) DetachGPU()      {}
func (h *Index) GPUFlush() error { return nil }
func (h *Index) GPUHealth() error { return nil }
func (h *Index) searchWithScoresGPU([]float32, int, *roaring.Bitmap, int) ([]types.SearchResult, bool) {
	return nil, false
}
func (h *Index) gpuNoteAdd(uint32, int)           {}
func (h *Index) gpuNoteRow(uint32, int, []uint32) {}
func (h *Index) gpuNoteDelete(uint32, bool)       {}
func (h *Index) gpuNoteRemove(uint32)             {}
func (h *Index) gpuNoteEntry()                    {}
