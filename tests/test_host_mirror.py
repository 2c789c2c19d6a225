"""Host-side C++ mirror of the reference interface (clstm_b200/host): registry, LCG init, `.clstm` proto2 codec.
The hand-written wire codec is cross-checked against python-protobuf with a descriptor built at run time from the
reference's schema (clstm.proto: KeyValue, Array, NetworkProto) -- no protoc needed."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "clstm_b200", "host")


def proto_classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    f = descriptor_pb2.FileDescriptorProto()
    f.name = "clstm_test.proto"
    f.package = "clstm"
    f.syntax = "proto2"
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name, fields):
        m = f.message_type.add()
        m.name = name
        for (fname, num, typ, label, tname) in fields:
            fd = m.field.add()
            fd.name, fd.number, fd.type, fd.label = fname, num, typ, label
            if tname:
                fd.type_name = tname
    REQ, OPT, REP = T.LABEL_REQUIRED, T.LABEL_OPTIONAL, T.LABEL_REPEATED
    msg("KeyValue", [("key", 1, T.TYPE_STRING, REQ, None), ("value", 2, T.TYPE_STRING, REQ, None)])
    msg("Array", [("name", 1, T.TYPE_STRING, OPT, None), ("dim", 2, T.TYPE_INT32, REP, None),
                  ("value", 3, T.TYPE_FLOAT, REP, None)])
    msg("NetworkProto", [("kind", 1, T.TYPE_STRING, REQ, None), ("name", 2, T.TYPE_STRING, OPT, None),
                         ("ninput", 10, T.TYPE_INT32, REQ, None), ("noutput", 11, T.TYPE_INT32, REQ, None),
                         ("icodec", 12, T.TYPE_INT32, REP, None), ("codec", 13, T.TYPE_INT32, REP, None),
                         ("attribute", 20, T.TYPE_MESSAGE, REP, ".clstm.KeyValue"),
                         ("weights", 30, T.TYPE_MESSAGE, REP, ".clstm.Array"),
                         ("sub", 40, T.TYPE_MESSAGE, REP, ".clstm.NetworkProto")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(f)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("clstm.NetworkProto"))


@pytest.fixture(scope="module")
def host_bin():
    import clstm_b200
    if not os.path.exists(clstm_b200.LIB_PATH):
        clstm_b200.build()
    subprocess.check_call(["make", "-s", "-C", HOST])
    return os.path.join(HOST, "test_host")


def flat_of_proto(net):
    """walk_params order, column-major flatten (what get_params returns)."""
    out = []

    def walk(n):
        for w in sorted(n.weights, key=lambda a: a.name):
            m = np.array(w.value, np.float32).reshape(w.dim[0], w.dim[1])   # stored row-major
            out.append(m.T.ravel())
        for s in n.sub:
            walk(s)
    walk(net)
    return np.concatenate(out)


def test_cpp_writer_matches_reference_schema_and_init(host_bin, tmp_path):
    from clstm_b200 import synth
    fn = str(tmp_path / "net.clstm")
    env = dict(os.environ, seed="0.222", EXPECT_NO_GPU="1")
    import torch
    if torch.cuda.is_available():
        env.pop("EXPECT_NO_GPU")
    out = subprocess.check_output([host_bin, "cpu", fn], env=env, text=True)
    assert "host cpu ok" in out
    Net = proto_classes()
    net = Net()
    net.ParseFromString(open(fn, "rb").read())
    assert net.kind == "Stacked" and net.ninput == 48 and net.noutput == 7
    assert [s.kind for s in net.sub] == ["Parallel", "SoftmaxLayer"]
    par = net.sub[0]
    assert [s.kind for s in par.sub] == ["NPLSTM", "Reversed"] and par.sub[1].sub[0].kind == "NPLSTM"
    assert sorted(w.name for w in par.sub[0].weights) == ["WCI", "WGF", "WGI", "WGO"]
    assert list(net.codec) == [0, 32, 97, 98, 99, 100, 101]
    attrs = {a.key: a.value for a in net.attribute}
    assert attrs["kind"] == "bidi" and attrs["trial"] == "1234" and "ninput" not in attrs
    # the host mirror's LCG init == the reference init restated in python (bit exact)
    assert np.array_equal(flat_of_proto(net), synth.reference_init(48, 10, 7, seed=0.222))


def test_cpp_reader_accepts_python_protobuf_files(host_bin, tmp_path):
    Net = proto_classes()
    rng = np.random.default_rng(0)
    ni, nh, nc = 6, 3, 4

    def lstm():
        n = Net(kind="NPLSTM", ninput=ni, noutput=nh)
        for name in ("WGI", "WGF", "WGO", "WCI"):
            w = n.weights.add()
            w.name = name
            w.dim.extend([nh, 1 + ni + nh])
            w.value.extend(rng.standard_normal(nh * (1 + ni + nh)).astype(np.float32).tolist())
        return n
    root = Net(kind="Stacked", ninput=ni, noutput=nc)
    root.codec.extend([0, 65, 66, 67])
    kv = root.attribute.add(); kv.key = "learning_rate"; kv.value = "0.001"
    par = root.sub.add(); par.kind = "Parallel"; par.ninput = ni; par.noutput = 2 * nh
    par.sub.append(lstm())
    rev = par.sub.add(); rev.kind = "Reversed"; rev.ninput = ni; rev.noutput = ni
    rev.sub.append(lstm())
    sm = root.sub.add(); sm.kind = "SoftmaxLayer"; sm.ninput = 2 * nh; sm.noutput = nc
    w = sm.weights.add(); w.name = "W1"; w.dim.extend([nc, 1 + 2 * nh])
    w.value.extend(rng.standard_normal(nc * (1 + 2 * nh)).astype(np.float32).tolist())
    fn = str(tmp_path / "py.clstm")
    open(fn, "wb").write(root.SerializeToString())
    out = subprocess.check_output([host_bin, "load", fn], text=True)
    flat = flat_of_proto(root)
    head = out.splitlines()[0]
    assert "kind=Stacked" in head and "ninput=%d" % ni in head and "noutput=%d" % nc in head
    assert "nparams=%d" % flat.size in head and "codec=4" in head and "lr=0.001" in head
    got = [float(l.split("=")[1]) for l in out.splitlines()[1:6]]
    assert np.allclose(got, flat[:5], rtol=1e-6)
    s = float(head.split("sum=")[1].split()[0])
    assert abs(s - float(flat.astype(np.float64).sum())) < 1e-4


@pytest.mark.gpu
def test_host_mirror_trains_and_round_trips_on_gpu(host_bin):
    out = subprocess.check_output([host_bin, "gpu"], text=True, timeout=300)
    assert "host gpu ok" in out
