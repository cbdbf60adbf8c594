"""CPU, build container only (needs /root/reference): the day files written by
meteor_scatter_b200.csvout are consumed by the reference's OWN dashboard reader
(database.py:61-106 load_last_30_days_csv_files, :242-287 scan_folder /
check_missing_days), i.e. the drop-in holds at the file contract."""
import datetime
import os
import sys
from unittest import mock

import numpy as np
import pytest

from meteor_scatter_b200 import csvout

REF = os.environ.get("MS_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "database.py")),
                                reason="reference checkout not available on this machine")


@pytest.fixture()
def ref_database(tmp_path, monkeypatch):
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.dates", "flask", "flask_apscheduler", "plotly",
                 "plotly.graph_objects", "plotly.io"):       # GUI / web imports of config.py:5-22, not used here
        monkeypatch.setitem(sys.modules, name, mock.MagicMock(name=name))
    monkeypatch.chdir(tmp_path)                     # config.py writes app.log into the cwd
    monkeypatch.syspath_prepend(REF)
    for m in ("database", "config"):
        sys.modules.pop(m, None)
    import database
    yield database
    for m in ("database", "config"):
        sys.modules.pop(m, None)


def test_reference_dashboard_reads_our_day_files(tmp_path, ref_database, capsys):
    folder = tmp_path / "csv-out"
    folder.mkdir()
    today = datetime.datetime.now().replace(hour=0, minute=0, second=0, microsecond=0)
    hour0 = today - datetime.timedelta(days=31)          # 30 full days up to yesterday (config.py:84-89)
    n_hours = 31 * 24
    rng = np.random.default_rng(0)
    hist = np.stack([rng.integers(0, 400, n_hours), rng.integers(0, 100, n_hours)], axis=1)
    hist[:, 1] = np.minimum(hist[:, 1], hist[:, 0])
    files = csvout.write_day_files(str(folder), csvout.hourly_rows(hist, hour0))
    assert len(files) == 31 and all(len(os.path.basename(f)) == 12 for f in files)

    found = ref_database.scan_folder(str(folder))                     # database.py:242-258
    assert sorted(found) == sorted(os.path.basename(f) for f in files)
    assert ref_database.check_missing_days(found) == []               # database.py:261-287
    df = ref_database.load_last_30_days_csv_files(str(folder))        # database.py:61-106
    assert list(df.columns) == ["Timestamp", "Anzahl", "Kritisch"]
    assert len(df) == n_hours
    assert int(df["Anzahl"].sum()) == int(hist[:, 0].sum()) and int(df["Kritisch"].sum()) == int(hist[:, 1].sum())
    import pandas as pd
    ts = pd.to_datetime(df["Timestamp"])                              # database.py:139, plot.py:203-215
    assert ts.min() == hour0 and ts.max() == hour0 + datetime.timedelta(hours=n_hours - 1)
    # the merged store the dashboard keeps (database.py:16-58) round-trips too
    store = tmp_path / "final_dataframe.csv"
    got = ref_database.load_or_create_dataframe(str(store), str(folder))
    assert os.path.exists(store) and len(got) == n_hours
