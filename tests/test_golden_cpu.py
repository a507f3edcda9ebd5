"""CPU: the oracle reproduces the committed golden fixtures (tests/golden/hotpath_v1.npz) — guards the oracle and
the fixture file against drifting apart."""
from oracle import riptide_oracle as orc

import golden_cases


def test_oracle_matches_golden_fixtures():
    golden_cases.run(orc)


def test_oracle_matches_golden_fixtures_v2():
    golden_cases.run_v2(orc)
