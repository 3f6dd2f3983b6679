"""CPU: Philox4x32-10 known-answer vectors (Random123 kat_vectors) for the oracle's stream port."""
from oracle.philox_port import philox4x32_10, PhiloxStream


def test_philox_known_answers():
    assert philox4x32_10([0, 0, 0, 0], (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox4x32_10([0xffffffff] * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_stream_uniforms_are_in_open_unit_interval():
    rs = PhiloxStream(1, 2, 3, 2)
    u = [rs.uniform() for _ in range(1000)]
    assert 0.0 < min(u) and max(u) < 1.0
    assert abs(sum(u) / 1000 - 0.5) < 0.05
