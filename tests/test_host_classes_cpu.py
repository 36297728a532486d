"""
The host layer above the C ABI (drop-in classes, simprint scoring, INSTANCE prefix index) on the oracle-backed store
double: the same test bodies the GPU suite runs through libisx_b200.so, executed here without a GPU so that the CPU
suite covers the host logic (argument handling, return shapes, float conversion, grouping, scoring, snapshots).
"""

import inspect

import pytest

from tests import test_gpu_instance, test_gpu_simprint, test_gpu_store

CASES = [
    test_gpu_store.test_nphd_index_add_get_contains_remove_semantics,
    test_gpu_store.test_nphd_search_return_shapes_and_float_conversion,
    test_gpu_store.test_index128_composite_keys_threshold_and_equality_modes,
    test_gpu_store.test_random_mutations_track_the_dict_model,
    test_gpu_store.test_snapshot_round_trip_and_reload,
    test_gpu_simprint.test_search_raw_matches_reference_fixtures,
    test_gpu_simprint.test_search_raw_with_index_doc_frequencies_equals_callback_path,
    test_gpu_simprint.test_search_exact_and_doc_freq_match_reference_fixtures,
    test_gpu_simprint.test_reference_behaviours_restated,
    test_gpu_simprint.test_first_of_asset_flags_equal_host_grouping,
    test_gpu_instance.test_instance_bidirectional_prefix_matches_reference_rules,
]


@pytest.mark.parametrize("case", CASES, ids=lambda f: f.__name__)
def test_host_layer_on_the_store_double(case, cpu_stores, tmp_path):
    kwargs = {"cuda": True}
    if "tmp_path" in inspect.signature(case).parameters:
        kwargs["tmp_path"] = tmp_path
    case(**kwargs)
