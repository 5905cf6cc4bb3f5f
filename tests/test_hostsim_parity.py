"""CPU-only parity tests: the device code compiled as a single-lane simulator (tests/hostsim) against the committed
golden GAF of the unmodified reference, and against oracle/_ref when it is present. Exercises host logic (GFA loader,
index build, batch dispatcher, result assembly, GAF writer) and the control flow of every kernel stage."""
import pytest

import cases
import mgtest as T


@pytest.fixture(scope="module")
def lib():
    return T.load_hostsim()


def test_c1_fixture_reads(lib, workdir):
    cases.case_c1(lib, workdir)


def test_c2_mt_synthetic(lib, workdir):
    cases.case_c2(lib, workdir)


def test_c3_sv_graph(lib, workdir):
    cases.case_c3(lib, workdir)


def test_c4_asm_preset(lib, workdir):
    cases.case_c4(lib, workdir)


def test_edge_reads(lib, workdir):
    cases.case_edge(lib, workdir)


def test_learned_tier_routing_keeps_results(lib, workdir):
    cases.case_tier_routing(lib, workdir)


def test_engine_switches_keep_results(lib, workdir):
    cases.case_switches(lib, workdir, device=False)


@pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built")
def test_struct_fields_vs_reference(lib, workdir):
    cases.case_struct_random(lib, workdir, n_reads=60)


@pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built")
def test_wfa_iteration_cap_fallback(lib):
    cases.case_wfa_fallback(lib)
