"""Length-prefixed protobuf framing — mirror of /root/reference/pkg/crowdllama/pbwire.go:14-70.

4-byte big-endian length, then the serialised BaseMessage; the read side rejects > 10 MiB."""
from __future__ import annotations

import struct

from .pb import BaseMessage

MAX_READ = 10 * 1024 * 1024   # pbwire.go:53


def write_length_prefixed_pb(w, msg: BaseMessage) -> None:
    data = msg.encode()
    if len(data) > 0xFFFFFFFF:                       # pbwire.go:22-26
        raise ValueError(f"message too large: {len(data)} bytes")
    try:
        w.write(struct.pack(">I", len(data)))
    except Exception as ex:                          # pbwire.go:32-34
        raise IOError(f"failed to write length prefix: {ex}") from ex
    try:
        w.write(data)
    except Exception as ex:                          # pbwire.go:37-39
        raise IOError(f"failed to write protobuf data: {ex}") from ex


def _read_full(r, n: int) -> bytes:
    buf = b""
    while len(buf) < n:
        chunk = r.read(n - len(buf))
        if not chunk:
            raise EOFError("unexpected EOF")
        buf += chunk
    return buf


def read_length_prefixed_pb(r) -> BaseMessage:
    try:
        hdr = _read_full(r, 4)
    except Exception as ex:                          # pbwire.go:47-49
        raise IOError(f"failed to read length prefix: {ex}") from ex
    (length,) = struct.unpack(">I", hdr)
    if length > MAX_READ:                            # pbwire.go:53-55
        raise ValueError(f"message too large: {length} bytes")
    try:
        data = _read_full(r, length)
    except Exception as ex:                          # pbwire.go:59-61
        raise IOError(f"failed to read protobuf data: {ex}") from ex
    try:
        return BaseMessage.decode(data)
    except Exception as ex:                          # pbwire.go:65-67
        raise ValueError(f"failed to unmarshal protobuf message: {ex}") from ex
