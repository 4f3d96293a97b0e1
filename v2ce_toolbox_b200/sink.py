"""Event-stream sink: move a large packed event stream from the device into one host array.

``tensor.cpu()`` on a multi-GB tensor runs at ~2 GB/s here: the driver stages pageable copies through small
internal buffers and every destination page is first-touched by that single copy thread.  For a 9000-frame clip
(21 GB of 13-byte records) that is ten seconds behind 0.3 s of device work.  ``to_host`` pipelines the transfer instead:
two pinned staging buffers filled by asynchronous D2H copies on a side stream, drained by a small thread pool whose
``numpy.copyto`` calls (GIL released) first-touch and fill disjoint slices of the destination in parallel.
"""
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

_staging = {}
_lock = threading.Lock()


def _buffers(chunk_bytes):
    with _lock:
        b = _staging.get(chunk_bytes)
        if b is None:
            b = _staging[chunk_bytes] = [torch.empty(chunk_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        return b


def _advise_huge_pages(arr):
    """Ask for transparent huge pages under a freshly allocated destination (madvise(MADV_HUGEPAGE) on the 2 MB-aligned
    interior): the copy threads then take one page fault per 2 MB instead of per 4 KB while they first-touch it.  A hint
    only -- ignored where THP is off."""
    try:
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        addr, nbytes = arr.ctypes.data, arr.nbytes
        lo = (addr + (2 << 20) - 1) & ~((2 << 20) - 1)
        hi = (addr + nbytes) & ~((2 << 20) - 1)
        if hi > lo:
            libc.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(hi - lo), 14)      # MADV_HUGEPAGE
    except Exception:                                   # noqa: BLE001
        pass


def to_host(dev_u8, chunk_bytes=128 << 20, workers=8, out=None):
    """dev_u8: 1-D uint8 CUDA tensor -> numpy uint8 array with the same bytes (``out`` if given)."""
    n = dev_u8.numel()
    if out is not None:
        dst = out
    else:
        dst = np.empty(n, dtype=np.uint8)
        if n >= (64 << 20):
            _advise_huge_pages(dst)
    if n == 0:
        return dst
    if n <= chunk_bytes:
        dst[:] = dev_u8.cpu().numpy()
        return dst
    bufs = _buffers(chunk_bytes)
    stream = torch.cuda.Stream(device=dev_u8.device)
    stream.wait_stream(torch.cuda.current_stream(dev_u8.device))
    sub = max(1, chunk_bytes // workers)
    pending = [None, None]                      # futures still reading staging buffer i
    with ThreadPoolExecutor(max_workers=workers) as pool:
        for k, off in enumerate(range(0, n, chunk_bytes)):
            i = k & 1
            if pending[i] is not None:
                for f in pending[i]:
                    f.result()                  # the workers are done with this staging buffer
            m = min(chunk_bytes, n - off)
            with torch.cuda.stream(stream):
                bufs[i][:m].copy_(dev_u8[off:off + m], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
            ev.synchronize()
            src = bufs[i].numpy()
            pending[i] = [pool.submit(np.copyto, dst[off + a:off + min(a + sub, m)], src[a:min(a + sub, m)])
                          for a in range(0, m, sub)]
        for p in pending:
            if p is not None:
                for f in p:
                    f.result()
    return dst


# ------------------------------------------------------------------------------------------
# .npz writer for the event stream (v2ce.py:371-372: np.savez(path, event_stream=...))
# ------------------------------------------------------------------------------------------
def _gf2_times(mat, vec):
    s, i = 0, 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_square(mat):
    return [_gf2_times(mat, m) for m in mat]


def crc32_combine(crc1, crc2, len2):
    """CRC-32 of A+B from crc32(A), crc32(B) and len(B) (zlib's crc32_combine: the CRC register is advanced over
    len2 zero bytes with GF(2) matrix squaring, so the cost is O(log len2))."""
    if len2 <= 0:
        return crc1
    odd = [0xEDB88320] + [1 << i for i in range(31)]      # operator for one zero bit
    even = _gf2_square(odd)                               # two zero bits
    odd = _gf2_square(even)                               # four zero bits
    while True:
        even = _gf2_square(odd)
        if len2 & 1:
            crc1 = _gf2_times(even, crc1)
        len2 >>= 1
        if not len2:
            break
        odd = _gf2_square(even)
        if len2 & 1:
            crc1 = _gf2_times(odd, crc1)
        len2 >>= 1
        if not len2:
            break
    return crc1 ^ crc2


def save_npz(path, chunk_bytes=32 << 20, workers=8, **arrays):
    """``np.savez(path, **arrays)`` for large arrays: the same uncompressed zip64 archive of ``<name>.npy`` members
    (``np.load`` reads it back unchanged), but the CRC-32 every zip member needs is computed chunk-wise on a thread pool
    (``zlib.crc32`` releases the GIL) while the main thread streams the bytes to the file, instead of serially in front of
    every write: ``np.savez`` spends ~0.55 s per GB on it, more than the device needs for the whole clip."""
    import io
    import os
    import struct
    import time
    import zlib

    path = str(path)
    if not path.endswith('.npz'):
        path += '.npz'
    tm = time.localtime()
    dos_time = (tm.tm_hour << 11) | (tm.tm_min << 5) | (tm.tm_sec // 2)
    dos_date = (max(tm.tm_year, 1980) - 1980) << 9 | (tm.tm_mon << 5) | tm.tm_mday
    central = []
    with open(path, 'wb') as f, ThreadPoolExecutor(max_workers=workers) as pool:
        for name, arr in arrays.items():
            arr = np.ascontiguousarray(arr)
            if arr.dtype.hasobject:
                raise TypeError('save_npz: object arrays are not supported')
            fname = (name + '.npy').encode()
            hdr = io.BytesIO()
            np.lib.format.write_array_header_1_0(hdr, np.lib.format.header_data_from_array_1_0(arr))
            hdr = hdr.getvalue()
            data = arr.reshape(-1).view(np.uint8) if arr.size else np.empty(0, np.uint8)
            size = len(hdr) + data.nbytes
            offset = f.tell()
            # local header: sizes in the zip64 extra field, like zipfile's force_zip64=True that np.savez uses
            extra = struct.pack('<HHQQ', 1, 16, size, size)
            f.write(struct.pack('<IHHHHHIIIHH', 0x04034b50, 45, 0, 0, dos_time, dos_date, 0, 0xFFFFFFFF, 0xFFFFFFFF,
                                len(fname), len(extra)) + fname + extra)
            f.write(hdr)
            f.flush()
            base = f.tell()
            fd = f.fileno()

            def put(a):
                # one worker per chunk: its CRC and its bytes, written in place (os.pwrite and zlib release the GIL),
                # so the page-cache copy of the member runs on all workers instead of one thread
                piece = data[a:a + chunk_bytes]
                done = 0
                while done < piece.nbytes:
                    done += os.pwrite(fd, piece[done:], base + a + done)
                return zlib.crc32(piece)

            futures = [pool.submit(put, a) for a in range(0, data.nbytes, chunk_bytes)]
            crc = zlib.crc32(hdr)
            for k, fut in enumerate(futures):
                crc = crc32_combine(crc, fut.result(), min(chunk_bytes, data.nbytes - k * chunk_bytes))
            f.seek(base + data.nbytes)
            end = f.tell()
            f.seek(offset + 14)
            f.write(struct.pack('<I', crc))
            f.seek(end)
            central.append((fname, crc, size, offset))
        cd_start = f.tell()
        for fname, crc, size, offset in central:
            big = b''
            if size >= 0xFFFFFFFF:
                big += struct.pack('<QQ', size, size)
            if offset >= 0xFFFFFFFF:
                big += struct.pack('<Q', offset)
            extra = struct.pack('<HH', 1, len(big)) + big if big else b''
            s32 = min(size, 0xFFFFFFFF)
            f.write(struct.pack('<IHHHHHHIIIHHHHHII', 0x02014b50, (3 << 8) | 45, 45, 0, 0, dos_time, dos_date, crc, s32, s32,
                                len(fname), len(extra), 0, 0, 0, 0o600 << 16, min(offset, 0xFFFFFFFF)) + fname + extra)
        cd_size = f.tell() - cd_start
        n = len(central)
        if cd_start >= 0xFFFFFFFF or cd_size >= 0xFFFFFFFF or n >= 0xFFFF:
            z64 = f.tell()
            f.write(struct.pack('<IQHHIIQQQQ', 0x06064b50, 44, 45, 45, 0, 0, n, n, cd_size, cd_start))
            f.write(struct.pack('<IIQI', 0x07064b50, 0, z64, 1))
        f.write(struct.pack('<IHHHHIIH', 0x06054b50, 0, 0, min(n, 0xFFFF), min(n, 0xFFFF), min(cd_size, 0xFFFFFFFF),
                            min(cd_start, 0xFFFFFFFF), 0))
    return path
