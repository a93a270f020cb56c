"""A minimal DICOM encoder for the tests (explicit / implicit VR little endian, one undefined-length sequence in front of
the pixel data): tests/test_io.py and tests/test_gpu_cli.py write their series with it."""
import struct


def _elem(tag, vr, value, explicit):
    if isinstance(value, str):
        value = value.encode("ascii")
        if len(value) % 2:
            value += b"\x00" if vr == "UI" else b" "
    g, e = tag
    if not explicit:
        return struct.pack("<HHI", g, e, len(value)) + value
    if vr in ("OB", "OW", "SQ", "UN"):
        return struct.pack("<HH2sHI", g, e, vr.encode(), 0, len(value)) + value
    return struct.pack("<HH2sH", g, e, vr.encode(), len(value)) + value


def _write_dicom(path, pixels, series, ipp, image_type="ORIGINAL\\PRIMARY\\AXIAL", explicit=True, slope=1, intercept=0,
                 iop="1\\0\\0\\0\\1\\0", signed=True, study="1.2.3", with_sequence=True):
    syntax = "1.2.840.10008.1.2.1" if explicit else "1.2.840.10008.1.2"
    meta = _elem((0x0002, 0x0010), "UI", syntax, True)
    meta = _elem((0x0002, 0x0000), "UL", struct.pack("<I", len(meta)), True) + meta
    body = b""
    body += _elem((0x0008, 0x0008), "CS", image_type, explicit)
    if with_sequence:   # an undefined-length sequence with one undefined-length item holding one element
        inner = _elem((0x0008, 0x0100), "SH", "CODE", explicit)
        item = struct.pack("<HHI", 0xFFFE, 0xE000, 0xFFFFFFFF) + inner + struct.pack("<HHI", 0xFFFE, 0xE00D, 0)
        seq = item + struct.pack("<HHI", 0xFFFE, 0xE0DD, 0)
        body += (struct.pack("<HH2sHI", 0x0008, 0x1140, b"SQ", 0, 0xFFFFFFFF) if explicit
                 else struct.pack("<HHI", 0x0008, 0x1140, 0xFFFFFFFF)) + seq
    body += _elem((0x0020, 0x000D), "UI", study, explicit)
    body += _elem((0x0020, 0x000E), "UI", series, explicit)
    body += _elem((0x0020, 0x0032), "DS", "\\".join("%g" % v for v in ipp), explicit)
    body += _elem((0x0020, 0x0037), "DS", iop, explicit)
    body += _elem((0x0028, 0x0002), "US", struct.pack("<H", 1), explicit)
    body += _elem((0x0028, 0x0010), "US", struct.pack("<H", pixels.shape[0]), explicit)
    body += _elem((0x0028, 0x0011), "US", struct.pack("<H", pixels.shape[1]), explicit)
    body += _elem((0x0028, 0x0030), "DS", "0.5\\0.75", explicit)
    body += _elem((0x0028, 0x0100), "US", struct.pack("<H", 16), explicit)
    body += _elem((0x0028, 0x0103), "US", struct.pack("<H", 1 if signed else 0), explicit)
    body += _elem((0x0028, 0x1052), "DS", "%g" % intercept, explicit)
    body += _elem((0x0028, 0x1053), "DS", "%g" % slope, explicit)
    body += _elem((0x7FE0, 0x0010), "OW", pixels.astype("<i2" if signed else "<u2").tobytes(), explicit)
    with open(path, "wb") as f:
        f.write(b"\x00" * 128 + b"DICM" + meta + body)
