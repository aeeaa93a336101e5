// Package b200handler is the cgo shim that drops the B200 engine into CrowdLlama's worker path.
//
// It builds a crowdllama.UnifiedAPIHandler (pkg/crowdllama/api.go:19) with the same contract as
// crowdllama.WorkerAPIHandler (api.go:45-96), replacing the HTTP round trip to Ollama
// (callOllamaAPI, api.go:108-160) by one blocking C call into libclengine.so.
//
// NOTE: this file cannot be compiled in the build image (no Go toolchain, no module proxy); it is
// the reference-side binding a maintainer adds.  The C-ABI it binds is exercised by the Python
// ctypes binding and tests in this repository.  See INTEGRATION.md.
package b200handler

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../crowdllama_b200/lib -lclengine -Wl,-rpath,${SRCDIR}/../../crowdllama_b200/lib
#include <stdlib.h>
#include "clengine.h"
extern int goFrameCallback(void* user, uint8_t* msg, size_t len);   // //export in callbacks.go
*/
import "C"

import (
	"context"
	"fmt"
	"runtime/cgo"
	"time"
	"unsafe"

	"google.golang.org/protobuf/proto"
	"google.golang.org/protobuf/types/known/timestamppb"

	llamav1 "github.com/crowdllama/crowdllama-pb/llama/v1"
	"github.com/crowdllama/crowdllama/pkg/crowdllama"
)

// Engine owns one GPU (one worker process per GPU; the 8 GPUs of a box are 8 worker peers).
type Engine struct{ h *C.cl_engine }

// Config mirrors the CROWDLLAMA_* knobs added for the B200 worker (pkg/config/config.go:58-79 style).
type Config struct {
	Device    int    // CROWDLLAMA_B200_DEVICE
	ModelName string // name advertised in Resource.SupportedModels, e.g. "llama3:8b"
	Preset    string // "llama3-8b" | "mistral-7b" | "tinyllama-1.1b"
	Seed      uint64
	MaxBatch  int
	KVBytes   int64
}

func lastErr(rc C.int) error {
	return fmt.Errorf("%s: %s", C.GoString(C.cl_strerror(rc)), C.GoString(C.cl_last_error()))
}

// New replaces "start the embedded Ollama server" (cmd/crowdllama/main.go:283-297).
func New(cfg Config) (*Engine, error) {
	var c C.cl_engine_config
	C.cl_default_engine_config(&c)
	name, preset := C.CString(cfg.ModelName), C.CString(cfg.Preset)
	defer C.free(unsafe.Pointer(name))
	defer C.free(unsafe.Pointer(preset))
	c.device, c.model_name, c.preset = C.int32_t(cfg.Device), name, preset
	c.weights_seed, c.max_batch, c.max_seqs = C.uint64_t(cfg.Seed), C.int32_t(cfg.MaxBatch), C.int32_t(cfg.MaxBatch)
	c.kv_pool_bytes, c.start_scheduler = C.int64_t(cfg.KVBytes), 1
	var h *C.cl_engine
	if rc := C.cl_engine_create(&c, &h); rc != C.CL_OK {
		return nil, lastErr(rc)
	}
	return &Engine{h: h}, nil
}

func (e *Engine) Close() { C.cl_engine_destroy(e.h) }

// Handler returns the drop-in for crowdllama.WorkerAPIHandler(ollamaBaseURL).  It is safe to call
// from many goroutines (one per inbound stream, pkg/peer/peer.go:177-182): cl_generate enqueues into
// the engine's continuous-batching scheduler and blocks; each in-flight call pins one OS thread.
func (e *Engine) Handler() crowdllama.UnifiedAPIHandler {
	return func(_ context.Context, req *llamav1.BaseMessage) (*llamav1.BaseMessage, error) {
		generateReq := req.GetGenerateRequest()
		if generateReq == nil { // api.go:48-51
			return nil, fmt.Errorf("expected GenerateRequest, got different message type")
		}
		model := C.CString(generateReq.Model)
		defer C.free(unsafe.Pointer(model))
		prompt := C.CString(generateReq.Prompt) // copied: C never retains Go memory
		defer C.free(unsafe.Pointer(prompt))
		var res C.cl_result
		// nil sampling = Ollama defaults (the reference sends no options, api.go:109-118)
		if rc := C.cl_generate(e.h, model, prompt, C.size_t(len(generateReq.Prompt)), nil, &res); rc != C.CL_OK {
			return nil, fmt.Errorf("failed to call B200 engine: %w", lastErr(rc))
		}
		defer C.cl_result_free(&res)
		return &llamav1.BaseMessage{Message: &llamav1.BaseMessage_GenerateResponse{
			GenerateResponse: &llamav1.GenerateResponse{
				Model:         generateReq.Model,
				CreatedAt:     timestamppb.Now(),
				Response:      C.GoStringN(res.text, C.int(res.text_len)),
				Done:          true,
				DoneReason:    C.GoString(res.done_reason),
				WorkerId:      "worker",              // api.go:83
				TotalDuration: time.Now().UnixNano(), // api.go:84 (kept bug-for-bug)
			}}}, nil
	}
}

// HandleStream is the streaming form (SURVEY.md §8f row 4; the reference rejects stream:true, api.go:155): the
// request is handed to libclengine as bytes, every response frame (Done=false text deltas, then Done=true) comes
// back through emit on the calling goroutine.  GenerateRequest.Options (field 4, the §8f-row-3 extension) is
// applied inside the library.  The frame callback is exported to C as goFrameCallback (see callbacks.go in a real
// build: //export goFrameCallback, cgo.Handle carries `emit`).
func (e *Engine) HandleStream(req *llamav1.BaseMessage, emit func(*llamav1.BaseMessage) error) error {
	raw, err := proto.Marshal(req)
	if err != nil {
		return err
	}
	h := cgo.NewHandle(func(frame []byte) error {
		var m llamav1.BaseMessage
		if err := proto.Unmarshal(frame, &m); err != nil {
			return err
		}
		return emit(&m)
	})
	defer h.Delete()
	buf := C.CBytes(raw) // copied: C never retains Go memory
	defer C.free(buf)
	if rc := C.cl_handle_message_stream(e.h, (*C.uint8_t)(buf), C.size_t(len(raw)), nil,
		C.cl_frame_cb(C.goFrameCallback), unsafe.Pointer(&h)); rc != C.CL_OK {
		return fmt.Errorf("failed to call B200 engine: %w", lastErr(rc))
	}
	return nil
}

// Stats feeds truthful routing metadata into crowdllama.Resource (pkg/crowdllama/types.go:30-40),
// replacing the constants at pkg/peer/peer.go:319-358.
func (e *Engine) Stats(r *crowdllama.Resource) error {
	var s C.cl_stats
	if rc := C.cl_engine_stats(e.h, &s); rc != C.CL_OK {
		return lastErr(rc)
	}
	r.TokensThroughput, r.Load = float64(s.tokens_per_sec), float64(s.load)
	r.VRAMGB, r.GPUModel = int(s.vram_gb), C.GoString(&s.gpu_model[0])
	return nil
}

// callbacks.go (same package; cgo requires //export functions to live in a file without C definitions):
//
//	//export goFrameCallback
//	func goFrameCallback(user unsafe.Pointer, msg *C.uint8_t, n C.size_t) C.int {
//		emit := (*cgo.Handle)(user).Value().(func([]byte) error)
//		if err := emit(C.GoBytes(unsafe.Pointer(msg), C.int(n))); err != nil {
//			return 1 // cancels the request: done_reason "cancelled"
//		}
//		return 0
//	}
