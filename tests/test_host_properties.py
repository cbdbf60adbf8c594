"""Property tests (hypothesis) of the host-side bookkeeping around the kernels: sharding, the hour grid, and the day
files the dashboard reads (`database.py:61-106`)."""
import datetime
import os

from hypothesis import given, settings, strategies as st

from meteor_scatter_b200 import csvout
from meteor_scatter_b200.batch import hour_span, shard_indices
from meteor_scatter_b200.pipeline import datetime_to_us, hour_index

naive_dt = st.datetimes(min_value=datetime.datetime(1960, 1, 1), max_value=datetime.datetime(2100, 1, 1))


@given(n=st.integers(0, 500), world=st.integers(1, 16))
def test_sharding_is_a_partition(n, world):
    parts = [shard_indices(n, r, world) for r in range(world)]
    assert sorted(i for p in parts for i in p) == list(range(n))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


@given(starts=st.lists(naive_dt, min_size=1, max_size=20), durs=st.lists(st.floats(0.0, 7200.0), min_size=20, max_size=20))
def test_hour_span_covers_every_recording(starts, durs):
    durs = durs[:len(starts)]
    hour0, n_hours = hour_span(starts, durs)
    assert hour0.minute == hour0.second == hour0.microsecond == 0 and n_hours >= 1
    end = hour0 + datetime.timedelta(hours=n_hours)
    for s, d in zip(starts, durs):
        assert hour0 <= s and s + datetime.timedelta(seconds=d) < end + datetime.timedelta(microseconds=1)
        # the hour index the kernel computes for an event at the very start of the recording lies inside the grid
        assert 0 <= hour_index(s) - hour_index(hour0) < n_hours


@given(dt=naive_dt)
def test_microsecond_clock_matches_python_floor_division(dt):
    us = datetime_to_us(dt)
    assert datetime.datetime(1970, 1, 1) + datetime.timedelta(microseconds=us) == dt
    assert hour_index(dt) == (dt.replace(minute=0, second=0, microsecond=0) - datetime.datetime(1970, 1, 1)) // \
        datetime.timedelta(hours=1)


@settings(max_examples=40, deadline=None)
@given(hour0=naive_dt.map(lambda d: d.replace(minute=0, second=0, microsecond=0)),
       counts=st.lists(st.tuples(st.integers(0, 500), st.integers(0, 500)), min_size=1, max_size=60),
       rerun=st.booleans())
def test_day_files_hold_one_row_per_hour_and_are_idempotent(tmp_path_factory, hour0, counts, rerun):
    folder = str(tmp_path_factory.mktemp("csv"))
    hist = [[a + k, k] for a, k in counts]                      # Kritisch <= Anzahl
    rows = csvout.hourly_rows(hist, hour0)
    written = csvout.write_day_files(folder, rows)
    if rerun:
        assert csvout.write_day_files(folder, rows) == written  # same batch again: nothing doubles
    seen = {}
    for path in written:
        name = os.path.basename(path)
        assert len(name) == 12 and name.endswith(".csv")         # database.py:77-81
        lines = open(path).read().splitlines()
        assert lines[0] == "Timestamp;Anzahl;Kritisch"
        stamps = [ln.split(";")[0] for ln in lines[1:]]
        assert stamps == sorted(stamps) and len(set(stamps)) == len(stamps)
        for ln in lines[1:]:
            ts, a, k = ln.split(";")
            t = datetime.datetime.strptime(ts, "%Y-%m-%d %H:%M:%S")
            assert t.strftime("%Y%m%d") + ".csv" == name
            seen[t] = (int(a), int(k))
    assert seen == {hour0 + datetime.timedelta(hours=i): (h[0], h[1]) for i, h in enumerate(hist)}
