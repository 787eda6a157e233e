# Warm worker of the host pool when the executor names its own interpreter
# (``python_path`` / ``conda_env``): started once with the reference's command line
#     [eval "$(conda shell.bash hook)" && conda activate <env> &&] <python_path> hostworker.py
# (/root/reference/covalent_ssh_plugin/ssh.py:377-380) and then fed electrons over its stdin,
# instead of paying an interpreter start per electron.  Semantics per task are exec.py's
# (exec.py:29-46): unpickle (fn, args, kwargs); create and enter the workdir; call; always return
# to the starting directory; ship (result, exception) back.  Self-contained on purpose — like
# exec.py it must run under ANY interpreter that has cloudpickle, without this package installed.
#
# Wire format, both directions: 8-byte little-endian length + pickle of
#     (blob, shm_path, [(offset, nbytes), ...])
# where blob is a protocol-5 pickle whose out-of-band buffers (large ndarrays) live in the
# shared-memory file shm_path (empty string: none) at the listed ranges — big payloads cross
# /dev/shm once instead of being copied through the pipe.
import mmap
import os
import pickle
import struct
import sys

OOB_MIN = 1 << 16            # buffers smaller than this stay in-band


def read_exact(stream, n):
    chunks = []
    while n:
        c = stream.read(n)
        if not c:
            raise EOFError
        chunks.append(c)
        n -= len(c)
    return b"".join(chunks)


def recv(stream):
    (n,) = struct.unpack("<Q", read_exact(stream, 8))
    return pickle.loads(read_exact(stream, n))


def send(stream, obj):
    raw = pickle.dumps(obj, protocol=4)
    stream.write(struct.pack("<Q", len(raw)))
    stream.write(raw)
    stream.flush()


def pack(codec, obj, shm_dir, tag):
    """obj -> (blob, shm_path, ranges): protocol 5 with large buffers in a /dev/shm file."""
    bufs = []

    def keep(b):
        if b.raw().nbytes >= OOB_MIN:
            bufs.append(b)
            return False          # out of band
        return True               # in band

    blob = codec.dumps(obj, protocol=5, buffer_callback=keep)
    if not bufs:
        return blob, "", []
    path = os.path.join(shm_dir, "b2pool-%d-%s" % (os.getpid(), tag))
    ranges, off = [], 0
    with open(path, "wb") as fh:
        for b in bufs:
            raw = b.raw()
            fh.write(raw)
            ranges.append((off, raw.nbytes))
            off += raw.nbytes
    return blob, path, ranges


def unpack(codec, blob, shm_path, ranges):
    if not shm_path:
        return codec.loads(blob), None
    with open(shm_path, "r+b") as fh:
        mm = mmap.mmap(fh.fileno(), 0)
    os.unlink(shm_path)           # the mapping keeps the pages alive for as long as the views do
    view = memoryview(mm)
    return codec.loads(blob, buffers=[view[o:o + n] for o, n in ranges]), mm


def main():
    inp, out = sys.stdin.buffer, sys.stdout.buffer
    sys.stdout = sys.stderr                       # electrons that print must not corrupt the channel
    try:
        import cloudpickle as codec
    except Exception as err:                      # exec.py:16-24: ship the import error, exit 1
        send(out, ("hello", repr(err)))
        sys.exit(1)
    send(out, ("hello", ""))
    shm_dir = "/dev/shm" if os.path.isdir("/dev/shm") else None
    serial = 0
    while True:
        try:
            blob, shm_path, ranges, workdir = recv(inp)
        except EOFError:
            return
        (fn, args, kwargs), _keep = unpack(codec, blob, shm_path, ranges)
        result, exception = None, None
        here = os.getcwd()
        try:
            os.makedirs(workdir or ".", exist_ok=True)
            os.chdir(workdir or ".")
            result = fn(*args, **kwargs)
        except Exception as err:                  # BaseException kills the worker, like a non-zero exit
            exception = err
        finally:
            os.chdir(here)
        serial += 1
        try:
            if shm_dir:
                reply = pack(codec, (result, exception), shm_dir, "r%d" % serial)
            else:
                reply = (codec.dumps((result, exception)), "", [])
        except Exception as err:
            reply = (codec.dumps((None, RuntimeError("result could not be pickled: %r" % (err,)))), "", [])
        send(out, reply)


if __name__ == "__main__":
    main()
