"""On-disk columnar segments (kb_segment_write / info / save, kb_store_append_file): the host-only half on the CPU, the round trip
through the device store with -m gpu."""
import os

import numpy as np
import pytest

from kolibrie_b200 import capi as c
from kolibrie_b200 import datagen
from tests import helpers as H
from tests import oracle_api as O


def test_segment_file_layout_host_only(tmp_path):
    d = datagen.employee_dataset(1000)
    path = str(tmp_path / "seg.kbs")
    c.segment_write(path, d.s, d.p, d.o, tag=42)
    n, tag, lo, hi = c.segment_info(path)
    assert (n, tag) == (d.n_triples, 42)
    assert lo == [int(d.s.min()), int(d.p.min()), int(d.o.min())] and hi == [int(d.s.max()), int(d.p.max()), int(d.o.max())]
    raw = np.fromfile(path, dtype=np.uint8)
    assert len(raw) % 4096 == 0 and len(raw) == 4096 + 3 * ((4 * n + 4095) // 4096 * 4096)
    col = lambda k: np.frombuffer(raw[4096 + k * ((4 * n + 4095) // 4096 * 4096):][: 4 * n].tobytes(), dtype=np.uint32)
    assert np.array_equal(col(0), d.s) and np.array_equal(col(1), d.p) and np.array_equal(col(2), d.o)
    (tmp_path / "junk").write_bytes(b"not a segment" * 400)
    with pytest.raises(c.KolibrieError):
        c.segment_info(str(tmp_path / "junk"))
    c.segment_write(str(tmp_path / "empty.kbs"), [], [], [], tag=1)
    assert c.segment_info(str(tmp_path / "empty.kbs"))[0] == 0


@pytest.mark.gpu
def test_segment_round_trip_through_the_device_store(ctx, tmp_path):
    d = datagen.employee_dataset(30000)
    js, pats, filt = datagen.employee_queries(d)["cfg2"]
    want = O.Db(d.s, d.p, d.o, d.num_or0, d.is_num).bgp(pats, filt)
    ctx.dict_numeric_load(d.num_or0, d.is_num)
    half = d.n_triples // 2 // 6 * 6
    ctx.store_clear()
    ctx.store_append(d.s[:half], d.p[:half], d.o[:half], tag=1)
    ctx.store_append(d.s[half:], d.p[half:], d.o[half:], tag=2)
    p_all, p_two = str(tmp_path / "all.kbs"), str(tmp_path / "two.kbs")
    ctx.segment_save(p_all)
    ctx.segment_save(p_two, tag=2)
    assert c.segment_info(p_all)[0] == d.n_triples and c.segment_info(p_two)[:2] == (d.n_triples - half, 2)
    # reload: whole store from one file
    ctx.store_clear()
    ctx.store_append_file(p_all, tag=7)
    s, p, o = ctx.store_download()
    assert np.array_equal(s, d.s) and np.array_equal(p, d.p) and np.array_equal(o, d.o)
    got = ctx.star_join(js, pats, filt)
    H.assert_same_bag(got.to_numpy(sorted(got.slots)), want.to_numpy(sorted(want.slots)), "reloaded store")
    # a window: first half resident and indexed, second half arrives from disk and the index is maintained
    ctx.store_clear()
    ctx.store_append(d.s[:half], d.p[:half], d.o[:half], tag=1)
    ctx.build_index()
    ctx.store_append_file(p_two, tag=2)
    n0 = ctx.get_stats()["index_joins"]
    got = ctx.star_join(js, pats, filt)
    assert ctx.get_stats()["index_joins"] == n0 + 1
    H.assert_same_bag(got.to_numpy(sorted(got.slots)), want.to_numpy(sorted(want.slots)), "segment appended from disk, index maintained")
    # a flipped byte is caught by the checksum; a truncated file by the read
    raw = bytearray(open(p_two, "rb").read())
    raw[4096 + 100] ^= 0x40
    open(str(tmp_path / "bad.kbs"), "wb").write(raw)
    n_before = ctx.store_size()[0]
    with pytest.raises(c.KolibrieError):
        ctx.store_append_file(str(tmp_path / "bad.kbs"), tag=3)
    open(str(tmp_path / "short.kbs"), "wb").write(raw[: len(raw) // 2])
    with pytest.raises(c.KolibrieError):
        ctx.store_append_file(str(tmp_path / "short.kbs"), tag=4)
    assert ctx.store_size()[0] == n_before, "a failed load leaves the store as it was"
