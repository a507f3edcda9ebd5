"""GPU: riptable_b200.rc reproduces the committed golden fixtures through the C ABI."""
import pytest

pytestmark = pytest.mark.gpu


def test_cuda_path_matches_golden_fixtures():
    import riptable_b200.rc as rc

    import golden_cases

    golden_cases.run(rc)


def test_cuda_path_matches_golden_fixtures_v2():
    import riptable_b200.rc as rc

    import golden_cases

    golden_cases.run_v2(rc)
