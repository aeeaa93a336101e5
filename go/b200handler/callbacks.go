package b200handler

// cgo requires //export functions to live in a file whose preamble holds declarations only.

/*
#include <stddef.h>
#include <stdint.h>
*/
import "C"

import (
	"runtime/cgo"
	"unsafe"
)

// goFrameCallback is the cl_frame_cb handed to cl_handle_message_stream (include/clengine.h): user points at the
// cgo.Handle of the Go closure that takes one serialised BaseMessage frame.  A nonzero return cancels the request
// inside the engine (done_reason "cancelled").
//
//export goFrameCallback
func goFrameCallback(user unsafe.Pointer, msg *C.uint8_t, n C.size_t) C.int {
	emit := (*cgo.Handle)(user).Value().(func([]byte) error)
	if err := emit(C.GoBytes(unsafe.Pointer(msg), C.int(n))); err != nil {
		return 1
	}
	return 0
}
