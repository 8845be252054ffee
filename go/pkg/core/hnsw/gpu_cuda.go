//go:build cuda

// Package hnsw — GPU mirror of the index behind the C ABI of libkektordb_gpu (include/kektordb_gpu.h).
//
// This file is what a KektorDB maintainer drops into pkg/core/hnsw to route SearchWithScores through a B200.
// It is selected by a build tag exactly like the Rust kernels (pkg/core/distance/distance_rust.go:1-17):
//
//	go build -tags cuda ./cmd/kektordb
//	CGO_CFLAGS="-I$KEKTOR_B200/include" CGO_LDFLAGS="-L$KEKTOR_B200/kektordb_b200 -lkektordb_gpu -lcudart -lstdc++"
//
// It cannot be compiled in the build image of the kernel repo (no Go toolchain there), so it is kept
// mechanical: every C call is a declared symbol of kektordb_gpu.h, every Go identifier it touches exists in
// the reference tree at the cited line.
//
// Shape (INTEGRATION.md §3): a request goroutine never blocks inside cgo.  It calls kdbgpu_batcher_submit
// (returns at once with a ticket), parks on a Go channel, and ONE dispatcher goroutine loops in
// kdbgpu_batcher_poll — the only OS thread that sits in C — waking the owners of finished tickets, which then
// copy their result out with kdbgpu_batcher_take.  Thousands of in-flight searches cost one OS thread plus the
// library's four batch workers, not one pinned thread each.
//
// The mirror follows the CPU index through a refresher (kdbgpu_refresher_*): Add / Delete / Vacuum / Refine
// report the rows they rewrote, and the library applies them in one exclusive section per max-lag /
// max-pending-rows (bounded staleness, INTEGRATION.md §4).
package hnsw

/*
#cgo CFLAGS: -I${SRCDIR}/../../../native/gpu/include
#cgo LDFLAGS: -lkektordb_gpu -lcudart -lstdc++
#include <stdlib.h>
#include "kektordb_gpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"log/slog"
	"path/filepath"
	"sync"
	"unsafe"

	"github.com/RoaringBitmap/roaring"
	"github.com/sanonone/kektordb/pkg/core/distance"
	"github.com/sanonone/kektordb/pkg/core/types"
)

// GPUOptions tunes the mirror.  Zero values pick the defaults in the comments.
type GPUOptions struct {
	Device         int    // CUDA device ordinal
	MaxBatch       uint32 // queries per device batch (1024)
	MaxWaitMicros  uint32 // micro-batcher deadline (1000)
	MaxPendingRows uint32 // refresher: flush when this many adjacency rows are queued (4096)
	MaxLagMillis   uint32 // refresher: flush when the oldest queued change is this old (50)
	CapacitySlack  uint32 // ids the mirror can hold beyond nodeCounter before it is re-attached (n/2 + 1024)
}

type gpuResult struct{}

type gpuMirror struct {
	h         *C.kdbgpu_index
	batcher   *C.kdbgpu_batcher
	refresher *C.kdbgpu_refresher
	dim       int
	capacity  uint32
	opt       GPUOptions // as resolved by AttachGPU: GPUHealth re-attaches with the same settings

	mu      sync.Mutex
	waiting map[uint64]chan gpuResult // ticket -> the goroutine parked on it
	early   map[uint64]struct{}       // tickets that finished before their owner registered
	filters map[*roaring.Bitmap]gpuFilter
	stop    chan struct{}
	done    chan struct{}
}

type gpuFilter struct {
	id   C.uint64_t
	card uint64
}

// gpuMirrors maps an index to its mirror without adding a field to Index (hnsw_index.go:40-134).
var gpuMirrors sync.Map // *Index -> *gpuMirror

func gpuErr(rc C.int) error {
	if rc == C.KDBGPU_OK {
		return nil
	}
	return errors.New(C.GoString(C.kdbgpu_last_error()))
}

func gpuOf(h *Index) *gpuMirror {
	if v, ok := gpuMirrors.Load(h); ok {
		return v.(*gpuMirror)
	}
	return nil
}

func gpuPrecision(p distance.PrecisionType) (C.int, error) {
	switch p { // distance.PrecisionType, pkg/core/distance/distance_go.go:41-46
	case distance.Float32:
		return C.KDBGPU_PRECISION_F32, nil
	case distance.Float16:
		return C.KDBGPU_PRECISION_F16, nil
	case distance.Int8:
		return C.KDBGPU_PRECISION_INT8, nil
	}
	return 0, fmt.Errorf("precision %v has no GPU mirror", p)
}

// AttachGPU mirrors the index into HBM and starts routing SearchWithScores through it.  Call it at the end of
// engine.Open (pkg/engine/engine.go:162-223), once the arena and the snapshot are loaded.
func (h *Index) AttachGPU(opt GPUOptions) error {
	if opt.MaxBatch == 0 {
		opt.MaxBatch = 1024
	}
	if opt.MaxWaitMicros == 0 {
		opt.MaxWaitMicros = 1000
	}
	if opt.MaxPendingRows == 0 {
		opt.MaxPendingRows = 4096
	}
	if opt.MaxLagMillis == 0 {
		opt.MaxLagMillis = 50
	}
	metric := C.int(C.KDBGPU_METRIC_L2)
	if h.metric == distance.Cosine { // hnsw_index.go:112
		metric = C.KDBGPU_METRIC_COSINE
	}
	prec, err := gpuPrecision(h.precision)
	if err != nil {
		return err
	}
	h.metaMu.RLock() // the same snapshot searchInternal takes (hnsw_index.go:373-382)
	n := uint32(h.nodeCounter.Load())
	entry := h.entrypointID.Load()
	maxLevel := int(h.maxLevel.Load())
	h.metaMu.RUnlock()
	if opt.CapacitySlack == 0 {
		opt.CapacitySlack = n/2 + 1024
	}
	g := &gpuMirror{dim: h.vectorDim, capacity: n + opt.CapacitySlack,
		waiting: map[uint64]chan gpuResult{}, early: map[uint64]struct{}{}, filters: map[*roaring.Bitmap]gpuFilter{},
		stop: make(chan struct{}), done: make(chan struct{})}
	g.opt = opt
	g.opt.CapacitySlack = 0 // sized again from the node count at re-attach time
	if err := gpuErr(C.kdbgpu_index_create_ex(C.int(opt.Device), C.int(h.vectorDim), metric, prec, C.int(h.m),
		C.uint32_t(g.capacity), &g.h)); err != nil {
		return err
	}
	fail := func(e error) error {
		C.kdbgpu_index_destroy(g.h)
		return e
	}
	if h.precision == distance.Int8 {
		if h.quantizer == nil { // hnsw_index.go:91, quantizer.go:19-22
			return fail(errors.New("int8 index without a quantizer"))
		}
		if err := gpuErr(C.kdbgpu_set_quantizer(g.h, C.float(h.quantizer.AbsMax))); err != nil {
			return fail(err)
		}
	}
	// ---- rows: straight from the arena's chunk files, placed at their logical ids by the library
	// (pkg/storage/mmap/arena.go:307-444); no copy through Go slices.
	if n > 0 {
		st := h.arena.GetState() // ArenaState.SlotTable: logical id -> physical slot (arena.go:252-268)
		dir := C.CString(h.arenaDir)
		var staged C.uint64_t
		rc := C.kdbgpu_arena_load_dir(g.h, dir, (*C.uint32_t)(unsafe.Pointer(&st.SlotTable[0])),
			C.uint32_t(len(st.SlotTable)), &staged)
		C.free(unsafe.Pointer(dir))
		if err := gpuErr(rc); err != nil {
			return fail(err)
		}
	}
	// ---- topology: the sidecar file when one is current, else flattened from the nodes once and written back
	sidecar := filepath.Join(h.arenaDir, "graph.kdbg")
	if err := g.stageGraph(h, sidecar, n, entry, maxLevel); err != nil {
		return fail(err)
	}
	if err := gpuErr(C.kdbgpu_prepare_search(g.h, C.uint32_t(opt.MaxBatch), 10, 128)); err != nil {
		return fail(err)
	}
	if err := gpuErr(C.kdbgpu_batcher_create(g.h, C.uint32_t(opt.MaxBatch), C.uint32_t(opt.MaxWaitMicros), &g.batcher)); err != nil {
		return fail(err)
	}
	if err := gpuErr(C.kdbgpu_refresher_create(g.h, C.uint32_t(opt.MaxPendingRows), C.uint32_t(opt.MaxLagMillis), &g.refresher)); err != nil {
		C.kdbgpu_batcher_destroy(g.batcher)
		return fail(err)
	}
	go g.dispatch()
	gpuMirrors.Store(h, g)
	return nil
}

// DetachGPU stops routing searches to the mirror and frees it.  Call it from Close (hnsw_index.go:3533) before
// the arena is unmapped.
func (h *Index) DetachGPU() {
	v, ok := gpuMirrors.LoadAndDelete(h)
	if !ok {
		return
	}
	g := v.(*gpuMirror)
	close(g.stop)
	<-g.done
	C.kdbgpu_refresher_destroy(g.refresher)
	C.kdbgpu_batcher_destroy(g.batcher)
	C.kdbgpu_index_destroy(g.h)
}

// stageGraph hands the topology to the library.  levels[i] = len(Connections)-1 (-1 = nil slot); node i owns
// rows nodeRow[i] .. nodeRow[i+1]-1, level 0 first (the layout of kdbgpu_set_graph).
func (g *gpuMirror) stageGraph(h *Index, sidecar string, n, entry uint32, maxLevel int) error {
	cs := C.CString(sidecar)
	defer C.free(unsafe.Pointer(cs))
	var fn, fentry C.uint32_t
	var fml C.int
	if C.kdbgpu_graph_file_probe(cs, &fn, nil, nil, nil, &fentry, &fml) == C.KDBGPU_OK &&
		uint32(fn) == n && uint32(fentry) == entry && int(fml) == maxLevel && !h.needsRefine.Load() {
		return gpuErr(C.kdbgpu_set_graph_file(g.h, cs)) // mapped and staged inside the library
	}
	nodes := h.getNodes() // hnsw_index.go:3634
	levels := make([]int32, n+1)
	nodeRow := make([]uint64, n+2)
	rowOff := []uint64{0}
	nbrs := make([]uint32, 0, int(n)*h.mMax0)
	deleted := make([]uint64, n/64+1)
	levels[0] = -1
	for id := uint32(1); id <= n; id++ {
		nodeRow[id] = uint64(len(rowOff) - 1)
		var nd *Node
		if int(id) < len(nodes) {
			nd = nodes[id]
		}
		if nd == nil {
			levels[id] = -1
			continue
		}
		h.RLockNode(id) // hnsw_index.go:2721
		levels[id] = int32(len(nd.Connections) - 1)
		for _, row := range nd.Connections {
			nbrs = append(nbrs, row...)
			rowOff = append(rowOff, uint64(len(nbrs)))
		}
		h.RUnlockNode(id)
		if nd.Deleted.Load() { // hnsw_node.go:37
			deleted[id/64] |= 1 << (id % 64)
		}
	}
	nodeRow[n+1] = uint64(len(rowOff) - 1)
	if len(nbrs) == 0 {
		nbrs = append(nbrs, 0)
	}
	if err := gpuErr(C.kdbgpu_set_graph(g.h, C.uint32_t(n), (*C.int32_t)(unsafe.Pointer(&levels[0])),
		(*C.uint64_t)(unsafe.Pointer(&nodeRow[0])), (*C.uint64_t)(unsafe.Pointer(&rowOff[0])),
		(*C.uint32_t)(unsafe.Pointer(&nbrs[0])), C.uint32_t(entry), C.int(maxLevel))); err != nil {
		return err
	}
	if err := gpuErr(C.kdbgpu_set_deleted(g.h, (*C.uint64_t)(unsafe.Pointer(&deleted[0])), C.size_t(len(deleted)))); err != nil {
		return err
	}
	// next start maps the file instead of walking the nodes
	if rc := C.kdbgpu_save_graph_file(g.h, cs); rc != C.KDBGPU_OK {
		slog.Warn("graph sidecar not written", "error", C.GoString(C.kdbgpu_last_error()))
	}
	return nil
}

// dispatch is the one goroutine that blocks inside C: it waits for finished tickets and wakes their owners.
func (g *gpuMirror) dispatch() {
	defer close(g.done)
	tickets := make([]C.uint64_t, 1024)
	for {
		select {
		case <-g.stop:
			return
		default:
		}
		var n C.uint32_t
		if C.kdbgpu_batcher_poll(g.batcher, &tickets[0], C.uint32_t(len(tickets)), 2000 /* µs */, &n) != C.KDBGPU_OK {
			return
		}
		if n == 0 {
			continue
		}
		g.mu.Lock()
		for i := 0; i < int(n); i++ {
			t := uint64(tickets[i])
			if ch, ok := g.waiting[t]; ok {
				delete(g.waiting, t)
				ch <- gpuResult{} // buffered: never blocks the dispatcher
			} else {
				g.early[t] = struct{}{} // finished before its owner parked
			}
		}
		g.mu.Unlock()
	}
}

// filterID returns the registered id of an allow-list, registering its dense form on first sight.  The engine
// builds one *roaring.Bitmap per FindIDsByFilter result (pkg/core/core.go:1766-1839) and passes the same pointer
// to every search of that request, so the pointer (guarded by the cardinality) keys the cache.
func (g *gpuMirror) filterID(allow *roaring.Bitmap, maxID uint32) (C.uint64_t, error) {
	card := allow.GetCardinality()
	g.mu.Lock()
	if f, ok := g.filters[allow]; ok && f.card == card {
		g.mu.Unlock()
		return f.id, nil
	}
	g.mu.Unlock()
	dense := make([]uint64, maxID/64+1) // same membership as the bitmap (hnsw_index.go:436-447, :2545-2549)
	it := allow.Iterator()
	for it.HasNext() {
		if id := it.Next(); id <= maxID {
			dense[id/64] |= 1 << (id % 64)
		}
	}
	var id C.uint64_t
	if err := gpuErr(C.kdbgpu_batcher_register_filter(g.batcher, (*C.uint64_t)(unsafe.Pointer(&dense[0])),
		C.size_t(len(dense)), &id)); err != nil {
		return 0, err
	}
	g.mu.Lock()
	if len(g.filters) >= 64 { // a small cache: release everything older
		for bm, f := range g.filters {
			C.kdbgpu_batcher_release_filter(g.batcher, f.id)
			delete(g.filters, bm)
		}
	}
	g.filters[allow] = gpuFilter{id: id, card: card}
	g.mu.Unlock()
	return id, nil
}

// gpuSearch is the body of the hook in SearchWithScores: one query in, its results out, batching underneath.
// efSearch already carries the needsRefine boost (hnsw_index.go:387-399); the library normalises for cosine
// exactly as normalize() does (:3030-3045).
func (g *gpuMirror) gpuSearch(query []float32, k, efSearch int, allow *roaring.Bitmap, maxID uint32) ([]types.SearchResult, error) {
	if len(query) != g.dim {
		return nil, fmt.Errorf("query has %d dimensions, index has %d", len(query), g.dim)
	}
	var filter C.uint64_t
	if allow != nil {
		if allow.IsEmpty() {
			return []types.SearchResult{}, nil // hnsw_index.go:443-445
		}
		var err error
		if filter, err = g.filterID(allow, maxID); err != nil {
			return nil, err
		}
	}
	var ticket C.uint64_t
	// the query is copied inside the call: no Go pointer is retained (cgo pointer rules)
	if err := gpuErr(C.kdbgpu_batcher_submit(g.batcher, (*C.float)(unsafe.Pointer(&query[0])), C.int(k), C.int(efSearch),
		nil, 0, filter, &ticket)); err != nil {
		return nil, err
	}
	t := uint64(ticket)
	g.mu.Lock()
	if _, ok := g.early[t]; ok {
		delete(g.early, t)
		g.mu.Unlock()
	} else {
		ch := make(chan gpuResult, 1)
		g.waiting[t] = ch
		g.mu.Unlock()
		<-ch // parked in Go, not in C
	}
	ids := make([]uint32, k)
	scores := make([]float64, k)
	var count C.uint32_t
	if err := gpuErr(C.kdbgpu_batcher_take(g.batcher, ticket, (*C.uint32_t)(unsafe.Pointer(&ids[0])),
		(*C.double)(unsafe.Pointer(&scores[0])), &count)); err != nil {
		return nil, err // SearchWithScores logs and returns [] (hnsw_index.go:355-359)
	}
	out := make([]types.SearchResult, int(count))
	for i := range out {
		out[i] = types.SearchResult{DocID: ids[i], Score: scores[i]} // Score = raw distance (:361-364)
	}
	return out, nil
}

// searchWithScoresGPU is called at the top of SearchWithScores (hnsw_index.go:343) after the closed checks:
//
//	if res, ok := h.searchWithScoresGPU(query, k, allowList, efSearch); ok {
//		return res
//	}
//
// ok == false means "no mirror attached": the CPU path below runs as before.  With a mirror attached an error
// is logged and answered with an empty slice, the reference's own convention (:355-359); there is no CPU
// fallback inside the library, and falling back here would hide a broken device behind a 50x slower path.
func (h *Index) searchWithScoresGPU(query []float32, k int, allowList *roaring.Bitmap, efSearch int) ([]types.SearchResult, bool) {
	g := gpuOf(h)
	if g == nil {
		return nil, false
	}
	if int(h.maxLevel.Load()) == -1 { // :383-385
		return []types.SearchResult{}, true
	}
	actualEf := efSearch
	if h.needsRefine.Load() { // :387-399
		boosted := int(float64(efSearch) * 2)
		if boosted < 80 {
			boosted = 80
		}
		if boosted > 200 {
			boosted = 200
		}
		if boosted > actualEf {
			actualEf = boosted
		}
	}
	res, err := g.gpuSearch(query, k, actualEf, allowList, uint32(h.nodeCounter.Load()))
	if err != nil {
		slog.Error("Error during GPU HNSW search", "error", err)
		return []types.SearchResult{}, true
	}
	return res, true
}

// ---- keeping the mirror fresh: called by the writers right after they release the node locks ----------------

// gpuNoteAdd: Add phase 1 (hnsw_index.go:559-655) — the node exists with its level and its stored vector.
func (h *Index) gpuNoteAdd(id uint32, level int) {
	g := gpuOf(h)
	if g == nil {
		return
	}
	raw, err := h.arena.GetBytes(id) // the stored form, whatever the precision (arena.go:378)
	if err != nil || len(raw) == 0 {
		return
	}
	C.kdbgpu_refresher_add_node(g.refresher, C.uint32_t(id), C.int(level), unsafe.Pointer(&raw[0]))
}

// gpuNoteRow: Connections[level] of node id was rewritten — forward / reverse links (:717-783), the batch commit
// (:1897-2060), reconnectNode (optimizer.go:195-222), Refine (:288-468).  row is read under the node's lock.
func (h *Index) gpuNoteRow(id uint32, level int, row []uint32) {
	g := gpuOf(h)
	if g == nil {
		return
	}
	var p *C.uint32_t
	if len(row) > 0 {
		p = (*C.uint32_t)(unsafe.Pointer(&row[0]))
	}
	C.kdbgpu_refresher_set_row(g.refresher, C.uint32_t(id), C.int(level), p, C.uint32_t(len(row)))
}

// gpuNoteDelete: Delete sets Node.Deleted (:2303-2336).
func (h *Index) gpuNoteDelete(id uint32, deleted bool) {
	if g := gpuOf(h); g != nil {
		d := C.int(0)
		if deleted {
			d = 1
		}
		C.kdbgpu_refresher_set_deleted(g.refresher, C.uint32_t(id), d)
	}
}

// gpuNoteRemove: Vacuum's physical cleanup, nodes[id] = nil (optimizer.go:252-274).
func (h *Index) gpuNoteRemove(id uint32) {
	if g := gpuOf(h); g != nil {
		C.kdbgpu_refresher_remove_node(g.refresher, C.uint32_t(id))
	}
}

// gpuNoteEntry: entrypointID / maxLevel moved (:793-801, optimizer.go:231-249).
func (h *Index) gpuNoteEntry() {
	if g := gpuOf(h); g != nil {
		C.kdbgpu_refresher_set_entry(g.refresher, C.uint32_t(h.entrypointID.Load()), C.int(h.maxLevel.Load()))
	}
}

// GPUFlush applies every queued change now (read-your-writes for callers that need it, e.g. tests).
func (h *Index) GPUFlush() error {
	if g := gpuOf(h); g != nil {
		if err := gpuErr(C.kdbgpu_refresher_flush(g.refresher)); err != nil {
			return errors.Join(err, h.GPUHealth())
		}
	}
	return nil
}

// GPUHealth re-stages the mirror from the CPU index when a flush has failed.  A flush that fails part-way (device out
// of memory, an id the mirror already holds) has consumed its batch (include/kektordb_gpu.h, refresher), so the mirror
// is behind the CPU index by those changes; background flushes report through stats.last_error only.  Call it from
// the maintenance ticker (optimizer.go:60-116) and after a failed GPUFlush.  nil = the mirror is current or absent.
func (h *Index) GPUHealth() error {
	g := gpuOf(h)
	if g == nil {
		return nil
	}
	var st C.kdbgpu_refresher_stats_t
	if C.kdbgpu_refresher_stats(g.refresher, &st) != C.KDBGPU_OK || st.last_error == C.KDBGPU_OK {
		return nil
	}
	slog.Error("GPU mirror refresh failed; re-staging from the CPU index", "code", int(st.last_error))
	opt := g.opt
	h.DetachGPU()
	return h.AttachGPU(opt)
}
