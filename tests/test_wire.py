"""Row f3 of SURVEY.md section 8: the reference's wire formats.  Messages are generated the way the reference's own code
emits them (oracle/wire_ref.py restates bin/sendStream.py:59-64 and bin/processStream.py:126-131); the expected values
are what ``json.loads`` returns for those strings.  CPU tests run the parser compiled for the host (the same source as
the device kernel); GPU tests decode whole triggers on the device and feed the ring buffers."""
import ctypes
import json
import math
import random
import struct

import numpy as np
import pytest
import torch

import tskd_b200
from tskd_b200 import capi
from tskd_b200 import stream as S
from conftest import load_golden
from oracle import stream_np as N
from oracle import wire_ref as R


def _record():
    g, _ = load_golden("p000194_replay.npz")
    return g, S.NumericsRecord(tuple(str(n) for n in g["names"]), g["gains"], g["baselines"], float(g["fs"]), g["raw"])


def _parse(s: bytes):
    st = ctypes.c_int32(0)
    return capi.load_library().b2cnn_parse_decimal(s, len(s), ctypes.byref(st)), st.value


def _same_bits(a: float, b: float) -> bool:
    return struct.pack("<d", a) == struct.pack("<d", b)


def test_host_parser_equals_json_loads_on_the_reference_messages_of_p000194():
    """Every [i, val] message sendStream.py would publish for the shipped record: val parsed == json.loads bit for bit."""
    g, rec = _record()
    sel = S.selected_signals(rec)
    names = [rec.names[i].replace(" ", "_") for i in sel]
    msgs = R.sample_messages(rec.physical[:, sel], names, "p000194-2112-05-23-14-34n")
    assert len(msgs) == 1625 * 4 and msgs[0][1] == b"p000194" and msgs[0][0] == "HR"
    n_nan = 0
    for topic, key, value in msgs:
        i, want = json.loads(value)
        body = value[value.index(b",") + 1:value.rindex(b"]")].strip()
        got, st = _parse(body)
        assert st == 0, value
        if isinstance(want, float) and math.isnan(want):
            assert math.isnan(got); n_nan += 1
        else:
            assert _same_bits(got, float(want)), value
    assert n_nan == int(np.isnan(rec.physical[:, sel]).sum())
    # a record with missing samples: json.dumps writes the bare token NaN (bin/sendStream.py does not filter it)
    rng = np.random.default_rng(3)
    p_signal = np.round(rng.normal(80, 20, size=(300, 5)), 1)
    p_signal[rng.random(p_signal.shape) < 0.3] = np.nan
    for topic, key, value in R.sample_messages(p_signal, list("abcde"), "p004980-x"):
        i, want = json.loads(value)
        got, st = _parse(value[value.index(b",") + 1:value.rindex(b"]")].strip())
        assert st == 0 and (math.isnan(got) if math.isnan(want) else _same_bits(got, want)), value


def test_host_parser_is_correctly_rounded():
    """Shortest-repr strings (Python repr, Java Double.toString), 17-digit cases, exact halfway cases, both exponent
    spellings -- against float(); out-of-range exponents and malformed text are flagged, never mis-parsed."""
    rng = random.Random(7)
    vals = [rng.uniform(0, 300) for _ in range(20000)] + [round(rng.uniform(0, 250), 1) for _ in range(5000)]
    vals += [rng.uniform(-1e-6, 1e-6) for _ in range(3000)] + [rng.uniform(-1e15, 1e15) for _ in range(3000)]
    vals += [rng.randint(0, 10 ** 17) / rng.choice([3, 7, 10, 1000]) for _ in range(5000)]
    for v in vals:
        for s in (repr(v), R.java_double_to_string(v)):
            got, st = _parse(s.encode())
            assert st == 0 and _same_bits(got, float(s)), s
    for s in ["0.30000000000000004", "9007199254740993", "9007199254740992", "4503599627370497.5", "4503599627370496.5",
              "1e22", "1e23", "8.5e-10", "123456789012345678", "1234567890123456789", "0.1", "1.0E-5", "1e-05", "-0.0",
              "2.5e+16", "0.000001", "1e27", "1E-27", "00012.50"]:
        got, st = _parse(s.encode())
        assert st == 0 and _same_bits(got, float(s)), s
    for s, want in [("NaN", math.nan), ('"NaN"', math.nan), ("null", math.nan), ("Infinity", math.inf), ("-Infinity", -math.inf)]:
        got, st = _parse(s.encode())
        assert st == 0 and (math.isnan(got) if math.isnan(want) else got == want)
    for s in ["abc", "1e", "--1", "", "1 ", "1,2", "0x10"]:
        assert _parse(s.encode())[1] == 1, s
    for s in ["1e28", "1e-300", "5e-324", "1.7976931348623157e308", "12345678901234567891"]:
        got, st = _parse(s.encode())
        assert st == 2 and math.isnan(got), s            # outside the supported range: flagged


def test_java_double_formatting_round_trips():
    for v in [81.0, 80.4, 1e-5, 1.5e-4, 0.001, 9999999.0, 1e7, 12345678.9, 100.0, 0.30000000000000004, 123456789012.0]:
        assert float(R.java_double_to_string(v)) == v
    assert R.java_double_to_string(1e-5) == "1.0E-5" and R.java_double_to_string(1e7) == "1.0E7"
    assert R.array_message("p000194", 3, [81.0, 80.4]) == (b"p000194_3", b"[81.0,80.4]")


def test_binary_frame_round_trip_and_validation():
    rng = np.random.default_rng(0)
    adc = rng.integers(-500, 3000, size=(3, 2, 7)).astype(np.int16)
    f = S.pack_frame([194, 195, 196], adc, first_index=120)
    assert len(f) == 32 + 12 + 4 + adc.nbytes
    ids, smp, first, grid = S.unpack_frame(f)
    assert list(ids) == [194, 195, 196] and np.array_equal(smp, adc) and first == 120 and not grid
    pts = rng.normal(80, 5, size=(2, 12, 10))
    ids, smp, first, grid = S.unpack_frame(S.pack_frame([1, 2], pts, first_index=7, grid_points=True))
    assert np.array_equal(smp, pts) and grid and first == 7
    with pytest.raises(RuntimeError):
        S.unpack_frame(f[:-2])                           # truncated
    with pytest.raises(RuntimeError):
        S.unpack_frame(b"XXXX" + f[4:])                  # bad magic
    with pytest.raises(ValueError):
        S.pack_frame([1], adc)


def test_binary_frame_header_counts_cannot_wrap_the_size_check():
    """An untrusted header whose counts multiply to 2^64 bytes (2^24 patients x 2^31 samples x 64 signals x 8 bytes)
    must not pass the length check of a 64 MB frame: the size is formed in 128-bit arithmetic."""
    lib = capi.load_library()
    n_pat, n_new, n_sig = 1 << 24, 1 << 31, 64
    head = struct.pack("<IHHIIIIQ", capi.FRAME_MAGIC, 1, capi.SAMPLES_F64, n_pat, n_new, n_sig, 0, 0)
    assert (8 * n_pat * n_new * n_sig) % (1 << 64) == 0
    frame = head + bytes(4 * n_pat)                       # header + ids: exactly the size a wrapped product describes
    buf = ctypes.create_string_buffer(frame, len(frame))
    hd, a, b = capi.FrameHeader(), ctypes.c_int64(0), ctypes.c_int64(0)
    rc = lib.b2cnn_frame_check(ctypes.addressof(buf), len(frame), ctypes.byref(hd), ctypes.byref(a), ctypes.byref(b))
    assert rc == capi.EINVAL and "header describes" in capi.last_error()


# ------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_decoder_rebuilds_the_physical_record_from_sendstream_messages():
    """All 6500 messages of p000194 in arrival order -> one device call -> the [1625, 4] fp64 frame == record.physical."""
    g, rec = _record()
    sel = S.selected_signals(rec)
    phys = rec.physical[:, sel]
    msgs = R.sample_messages(phys, [rec.names[i] for i in sel], "p000194-2112-05-23-14-34n")
    rows = np.repeat(np.arange(1625), 4)                 # message t belongs to sample t // 4 (arrival order)
    frame, bad = S.decode_sample_messages([m[2] for m in msgs], rows, 1625, 4, "cuda:0")
    got = frame.cpu().numpy()
    assert bad == 0 and np.array_equal(np.isnan(got), np.isnan(phys))
    assert np.array_equal(got[~np.isnan(phys)].view(np.int64), phys[~np.isnan(phys)].view(np.int64))     # bit for bit
    # malformed and out-of-range messages are counted, not mis-parsed; a signal without a message stays NaN
    frame, bad = S.decode_sample_messages([b"[0, 81.0]", b"[1, oops]", b"[2 81.0]", b"[3, 1e99]", b"[9, 5.0]"], [0, 0, 0, 0, 0], 1, 4, "cuda:0")
    assert bad == 3 and frame.cpu().numpy()[0, 0] == 81.0 and np.isnan(frame.cpu().numpy()[0, 1:]).all()
    # a row outside the frame (caller error) is dropped, not written: the frame is the tail of a larger buffer here
    frame, bad = S.decode_sample_messages([b"[0, 1.0]", b"[1, 2.0]", b"[2, 3.0]"], [0, 1, -1], 1, 4, "cuda:0")
    assert bad == 0 and frame.cpu().numpy()[0, 0] == 1.0 and np.isnan(frame.cpu().numpy()[0, 1:]).all()


@pytest.mark.gpu
def test_messages_to_ring_equals_whole_record_windows():
    """sendStream messages -> device decoder -> ring buffers (one trigger = one sample at 1/60 Hz) == b2cnn_prep_windows."""
    g, rec = _record()
    sel = S.selected_signals(rec)
    msgs = R.sample_messages(rec.physical[:, sel], [rec.names[i] for i in sel], "p000194-2112-05-23-14-34n")
    whole, _ = S.assemble_windows_gpu(rec, "cuda:0")
    ring = S.PatientRing(1, 4, rec.fs, device="cuda:0")
    ring.set_signals(0, [0, 1, 2, 3])
    out = []
    for i in range(400):                                 # 400 triggers: 391 windows
        frame, bad = S.decode_sample_messages([m[2] for m in msgs[4 * i:4 * i + 4]], [0, 0, 0, 0], 1, 4, "cuda:0")
        assert bad == 0
        r = ring.push(frame.view(1, 1, 4))
        if r is not None:
            out.append(r[0][0].clone())
    assert len(out) == 391 and torch.equal(torch.stack(out), whole[:391])


@pytest.mark.gpu
def test_call_stream_array_messages_to_ring():
    """processStream's call-stream payload (12 grid points per channel and trigger as a JSON array printed by the JVM)
    -> device decoder -> ring (grid points appended as they are) == the windows cut from the pandas-built grids."""
    g, rec = _record()
    grids = g["grids"]                                   # [4][19489] from pandas (make_golden.py)
    want, _ = N.windows_from_grids(grids)
    ring = S.PatientRing(2, 10, rec.fs, device="cuda:0")
    for p in range(2):
        ring.set_signals(p, [0, 1, 2, 3])
    out = []
    for trig in range(60):
        msgs = [R.array_message("p000194", c, grids[c, 12 * trig:12 * trig + 12]) for c in range(4)]
        vals, counts, bad = S.decode_array_messages([m[1] for m in msgs], 12, "cuda:0")
        assert bad == 0 and (counts == 12).all()
        assert np.array_equal(vals.cpu().numpy(), grids[:, 12 * trig:12 * trig + 12])           # exact doubles
        pts = torch.full((2, 12, 10), float("nan"), dtype=torch.float64, device="cuda:0")
        pts[:, :, :4] = vals.t().unsqueeze(0)            # channel index from the message key "<pid>_<chan>"
        r = ring.push(pts, grid_points=True)
        if r is not None:
            out.append(r[0][0].clone())
    assert len(out) == 51
    assert torch.equal(torch.stack(out).cpu(), torch.from_numpy(want[:51].astype(np.float32)))
    vals, counts, bad = S.decode_array_messages([b"[]", b"[1.0,2.0", b'["NaN",3.5]', b"[1.0,2.0,3.0]"], 2, "cuda:0")
    c = counts.cpu().numpy()
    assert c[0] == 0 and c[1] == -1 and c[2] == 2 and c[3] == 3 and bad == 2
    assert np.isnan(vals.cpu().numpy()[2, 0]) and vals.cpu().numpy()[2, 1] == 3.5


@pytest.mark.gpu
def test_binary_frames_feed_the_ring_without_decoding():
    g, rec = _record()
    whole, _ = S.assemble_windows_gpu(rec, "cuda:0")
    ring = S.PatientRing(2, 7, rec.fs, device="cuda:0")
    for p in range(2):
        ring.set_record_signals(p, rec)
    out = []
    for i in range(200):
        frame = S.pack_frame([194, 195], np.repeat(rec.raw[None, i:i + 1], 2, axis=0), first_index=i)
        ids, smp, first, grid = S.unpack_frame(frame)
        assert first == i and list(ids) == [194, 195]
        r = ring.push(smp)
        if r is not None:
            out.append(r[0].clone())
    assert len(out) == 191 and torch.equal(torch.stack([o[1] for o in out]), whole[:191])
