// b2_hostcopy.h — host-side copy pool: caller buffers -> pinned ingress memory.
//
// The replacement of "pickle the arguments into function_<id>.pkl"
// (/root/reference/covalent_ssh_plugin/ssh.py:147-150): operand bytes are gathered
// straight from the caller's arrays into the pinned staging buffer the next H2D reads,
// by a small pool of host threads using non-temporal stores (the destination is never
// read by the CPU again, so it should neither be read-for-ownership nor fill the caches).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace b2host {

struct CopyPiece {
  const uint8_t *src;
  uint8_t *dst;
  size_t bytes;
};

// Copies every piece; returns when all of them are done.  The calling thread works too.
// Thread count: B2_STAGE_THREADS (default: min(8, hardware threads)); 1 = caller only.
void copy_pieces(const CopyPiece *pieces, size_t n);

// One streaming copy on the calling thread.
void copy_stream(uint8_t *dst, const uint8_t *src, size_t bytes);

int pool_threads();

}  // namespace b2host
