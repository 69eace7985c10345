"""Tampered restart checkpoints: one mutation per rule of the reference's loader
(reference src/sipnet/restart.c:593-757 parse rules, :963-983 load-time checks; the same cases as its
tests/sipnet/test_restart_infrastructure/testRestartMVP.c).  tests/golden/make_golden.py runs the UNMODIFIED
reference binary on each and records its exit code in tests/golden/restart_cases.json; tests/test_restart.py
asks the host library the same question."""
import re


def _sub(pattern, repl, count=1):
    return lambda t: re.sub(pattern, repl, t, count=count, flags=re.M)


CASES = {
    "untouched": lambda t: t,
    "bad_magic": _sub(r"^SIPNET_RESTART$", "SIPNET_RESTORE"),
    "empty_file": lambda t: "",
    "duplicate_key": lambda t: t.replace("envi.snow ", "envi.snow 0\nenvi.snow ", 1),
    "unknown_key": lambda t: t.replace("end_restart 1", "envi.bogus 1\nend_restart 1"),
    "unknown_key_known_prefix": lambda t: t.replace("end_restart 1", "trackers.bogus 1\nend_restart 1"),
    "missing_envi_key": _sub(r"^envi\.snow .*\n", ""),
    "missing_tracker_key": _sub(r"^trackers\.meanNPP .*\n", ""),
    "missing_flag": _sub(r"^flags\.flooding .*\n", ""),
    "missing_end_marker": _sub(r"^end_restart 1\n", ""),
    "missing_ring_value": _sub(r"^mean\.npp\.values\.17 .*\n", ""),
    "missing_ring_weight": _sub(r"^mean\.npp\.weights\.249 .*\n", ""),
    "ring_index_out_of_range": lambda t: t.replace("end_restart 1", "mean.npp.values.250 0\nend_restart 1"),
    "ring_index_not_a_number": lambda t: t.replace("end_restart 1", "mean.npp.values.x 0\nend_restart 1"),
    "duplicate_ring_slot": lambda t: t.replace("end_restart 1", "mean.npp.weights.3 0\nend_restart 1"),
    "nan_value": _sub(r"^envi\.soilC .*$", "envi.soilC nan"),
    "inf_value": _sub(r"^trackers\.totNee .*$", "trackers.totNee inf"),
    "trailing_garbage_in_double": _sub(r"^envi\.soilC (.*)$", r"envi.soilC \1x"),
    "double_where_int": _sub(r"^phenology\.lastYear .*$", "phenology.lastYear 2016.5"),
    "int_overflow": _sub(r"^trackers\.lastYear .*$", "trackers.lastYear 4294967296"),
    "three_tokens": _sub(r"^envi\.snow (.*)$", r"envi.snow \1 extra"),
    "one_token": _sub(r"^envi\.snow .*$", "envi.snow"),
    "schema_mismatch": _sub(r"^schema_layout\.trackers_size .*$", "schema_layout.trackers_size 256"),
    "missing_schema_key": _sub(r"^schema_layout\.survival_trackers_size .*\n", ""),
    "flag_mismatch": _sub(r"^flags\.litterPool .*$", "flags.litterPool 0"),
    "version_mismatch": _sub(r"^meta_info\.model_version .*$", "meta_info.model_version 2.0.9"),
    "build_info_mismatch": _sub(r"^meta_info\.build_info .*$", "meta_info.build_info someone_else"),
    "ring_resized": _sub(r"^mean\.npp\.length .*$", "mean.npp.length 100"),
    "cursor_out_of_range": _sub(r"^mean\.npp\.start .*$", "mean.npp.start -1"),
    "boundary_after_segment_start": _sub(r"^boundary\.year .*$", "boundary.year 2017"),
    "boundary_zero_length": _sub(r"^boundary\.length .*$", "boundary.length 0"),
    "boundary_gap_warns_only": _sub(r"^boundary\.time .*$", "boundary.time 12"),
    "text_after_end_marker": lambda t: t + "anything at all here\n",
    "blank_and_indented_lines": lambda t: t.replace("envi.snow ", "\n   \n\t envi.snow ", 1),
    "crlf_header": lambda t: t.replace("SIPNET_RESTART\n", "SIPNET_RESTART\r\n", 1),
    "overlong_line": lambda t: t.replace("end_restart 1", "envi.padding " + "9" * 5000 + "\nend_restart 1"),
}
