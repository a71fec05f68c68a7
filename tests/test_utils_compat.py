"""CPU-only: the helper modules the reference's scripts import (utils.logger_utils, utils.info_collector_callback)."""
import logging
import os

import helpers as H  # noqa: F401  (sys.path)
from utils.info_collector_callback import InfoCollectorCallback
from utils.logger_utils import setup_logger


def test_setup_logger_writes_console_and_file(tmp_path):
    log = setup_logger("SampleRL_test", str(tmp_path), console_level=logging.WARNING, file_level=logging.DEBUG)
    log.debug("to the file only")
    assert setup_logger("SampleRL_test", str(tmp_path)) is log and len(log.handlers) == 2  # handlers are added once
    for h in log.handlers:
        h.flush()
    assert "to the file only" in open(os.path.join(str(tmp_path), "SampleRL_test.log")).read()


def test_info_collector_counts_outcomes_per_100_episodes(tmp_path):
    cb = InfoCollectorCallback()
    outcomes = ["Goal"] * 60 + ["Out"] * 30 + ["Timeout"] * 10 + ["Goal"] * 50
    for r in outcomes:
        cb.locals = {"infos": [{"result": r}, {"result": None}]}  # running episodes carry no result
        assert cb._on_step() is True
    assert len(cb.infos) == 150
    goal, out, timeout = cb.plot_print_results(logging.getLogger("x"), file_name=str(tmp_path / "train_results"))
    assert goal == [60.0, 100.0] and out == [30.0, 0.0] and timeout == [10.0, 0.0]
    assert os.path.exists(str(tmp_path / "train_results.png")) or os.path.exists(str(tmp_path / "train_results.csv"))
    cb.reset()
    assert cb.infos == [] and cb.results == {}
