"""The device-side f64 formatter (loro_b200/csrc/lb_f64.cuh: exact shortest round-trip digits, serde_json / ryu layout)
against the oracle's (std::to_chars shortest digits + the same layout rules), on the host through the emulated build."""
import ctypes
import os
import random
import struct
import subprocess

import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu", "libloro_b200_emu.so")


@pytest.fixture(scope="module")
def fmt():
    subprocess.check_call([os.path.join(HERE, "emu", "build_emu.sh")])
    L = ctypes.CDLL(EMU)
    L.lb_emu_format_f64.argtypes = [ctypes.c_double, ctypes.c_char_p]
    buf = ctypes.create_string_buffer(64)

    def f(d):
        n = L.lb_emu_format_f64(d, buf)
        return buf.raw[:n]
    return f


def want(d):
    return oracle.codec("f64_json", struct.pack("<d", d))


def test_known_values(fmt):
    cases = {0.0: b"0.0", 1.0: b"1.0", 0.1: b"0.1", 0.3: b"0.3", 1.5: b"1.5", 100.0: b"100.0", 1e16: b"1e16",
             1e15: b"1000000000000000.0", 123456789012345680.0: b"1.2345678901234568e17", 1e-5: b"0.00001", 1e-6: b"1e-6",
             5e-324: b"5e-324", 1.7976931348623157e308: b"1.7976931348623157e308", 2.2250738585072014e-308: b"2.2250738585072014e-308",
             3.141592653589793: b"3.141592653589793", -2.5e-7: b"-2.5e-7", 9007199254740993.0: b"9007199254740992.0",
             0.000123: b"0.000123", 4.35: b"4.35", 2.0 ** 60: b"1.152921504606847e18", float("inf"): b"null", float("nan"): b"null"}
    assert fmt(-0.0) == b"-0.0" and want(-0.0) == b"-0.0"
    for d, s in cases.items():
        assert fmt(d) == s, (d, fmt(d), s)
        assert want(d) == s, (d, want(d))


def test_random_doubles_match_the_oracle(fmt):
    rnd = random.Random(5)
    n = 0
    for _ in range(60000):
        k = rnd.random()
        if k < 0.4:
            d = struct.unpack("<d", struct.pack("<Q", rnd.getrandbits(64)))[0]
        elif k < 0.6:
            d = rnd.uniform(-1e6, 1e6)
        elif k < 0.7:
            d = round(rnd.uniform(-1000, 1000), rnd.randint(0, 6))
        elif k < 0.8:
            d = float(rnd.randint(-2 ** 62, 2 ** 62))
        elif k < 0.9:
            d = 2.0 ** rnd.randint(-1074, 1023) * rnd.choice([1, -1, 1.5, 3])
        else:
            d = rnd.choice([1, 5, 25, 125]) * 10.0 ** rnd.randint(-30, 30)
        if d != d or d in (float("inf"), float("-inf")):
            assert fmt(d) == b"null"
            continue
        got = fmt(d)
        assert got == want(d), (d.hex(), got, want(d))
        assert float(got) == d
        n += 1
    assert n > 50000
