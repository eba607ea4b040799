"""The teacher's .meta contract (model.py:122-134, 323-341) without TensorFlow: MetaGraphDef skeleton reader / writer."""
import numpy as np
import pytest

from sr_wavenet_b200 import tf_meta
from sr_wavenet_b200.nsynth import _ld, _enc_varint


def test_meta_roundtrip_and_contract(tmp_path):
    nodes, cols = tf_meta.teacher_meta_skeleton()
    p = str(tmp_path / 'model.ckpt-3.meta')
    tf_meta.write_meta(p, nodes, cols)
    m = tf_meta.read_meta(p)
    assert m['nodes'] == nodes and m['collections'] == cols
    picked = tf_meta.check_teacher_contract(m)
    assert picked['Logits_d'] == 'WaveNetAutoEncoder/Decoder_1/logits:0'        # the reuse=True decoder fed from the placeholder
    assert set(tf_meta.TEACHER_COLLECTIONS) == set(cols)


def test_meta_reader_on_a_hand_assembled_message(tmp_path):
    """Field numbers of meta_graph.proto / graph.proto / node_def.proto, other fields skipped: meta_info_def (1), a node with
    inputs and a device, saver_def (3), a bytes_list collection (trainable_variables) and a node_list collection."""
    node = _ld(1, b'scope/x') + _ld(2, b'Placeholder') + _ld(3, b'^dep') + _ld(4, b'/cpu:0')
    node2 = _ld(1, b'scope/y') + _ld(2, b'Identity') + _ld(3, b'scope/x')
    graph = _ld(1, node) + _ld(1, node2) + _ld(4, _enc_varint((1 << 3) | 0) + _enc_varint(27))       # versions { producer: 27 }
    tv = _ld(4, _ld(1, b'trainable_variables') + _ld(2, _ld(2, _ld(1, b'\x0a\x03abc') + _ld(1, b'\x0a\x03def'))))
    nl = _ld(4, _ld(1, b'Logits_d') + _ld(2, _ld(1, _ld(1, b'scope/y:0'))))
    msg = _ld(1, _ld(1, b'meta info')) + _ld(2, graph) + _ld(3, _ld(1, b'save/Const:0')) + tv + nl
    p = tmp_path / 'm.meta'
    p.write_bytes(msg)
    m = tf_meta.read_meta(str(p))
    assert m['nodes'] == {'scope/x': 'Placeholder', 'scope/y': 'Identity'}
    assert m['collections'] == {'trainable_variables': [None, None], 'Logits_d': ['scope/y:0']}


def test_contract_violations(tmp_path):
    nodes, cols = tf_meta.teacher_meta_skeleton()
    del cols['Logits_d']
    p = str(tmp_path / 'a.meta')
    tf_meta.write_meta(p, nodes, cols)
    with pytest.raises(IndexError):
        tf_meta.check_teacher_contract(tf_meta.read_meta(p))
    nodes, cols = tf_meta.teacher_meta_skeleton()
    del nodes['WaveNetAutoEncoder/encoding_nodecoder_placeholder']
    tf_meta.write_meta(p, nodes, cols)
    with pytest.raises(ValueError):
        tf_meta.check_teacher_contract(tf_meta.read_meta(p))
