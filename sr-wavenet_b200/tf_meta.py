"""The teacher's ``.meta`` contract (model.py:122-134 / 323-334) without TensorFlow.

``tf.train.Saver.save`` writes ``model.ckpt-N.meta`` next to the checkpoint: a serialized ``MetaGraphDef``.  The student
imports it (``tf.train.import_meta_graph(path, input_map=...)``, model.py:326-331), re-wires three of the teacher's
placeholders and picks the teacher's tensors out of the graph's COLLECTIONS (``Logits_d``, ``Encoding_output``,
``Inputs_e``, ``Out_e``, ``Out_d``; model.py:333-341).  What this module keeps of that contract:

* ``read_meta``: parses the protobuf wire format of a ``MetaGraphDef`` far enough to list the graph's nodes (name, op)
  and its collections (``collection_def``: ``node_list`` values, i.e. tensor names) -- field numbers from
  tensorflow/core/protobuf/meta_graph.proto and tensorflow/core/framework/{graph,node_def}.proto;
* ``write_meta``: writes a minimal ``MetaGraphDef`` with the same structure (placeholder nodes + collections), which is
  what ``WaveNetAutoEncoder.save`` puts next to its checkpoints;
* ``check_teacher_contract``: what ``import_meta_graph`` + the ``get_collection(...)[0]`` lookups of model.py:326-341 would
  fail on: an ``input_map`` key that is not a tensor of the graph, a missing collection.

The computation is not imported from the file (there is no graph executor here): the teacher is rebuilt from the
checkpoint's variables, and the collections map onto methods of ``WaveNetAutoEncoder`` (``TEACHER_COLLECTIONS``).
Parity unpinned against TensorFlow itself: no TF-written ``.meta`` exists in this environment; the parser is tested
against hand-assembled messages and its own writer (tests/test_tf_meta.py).
"""
from .nsynth import _fields, _ld

# collection (model.py:122-134) -> what provides it here
TEACHER_COLLECTIONS = {
    'Inputs_e': 'inputs of encode() / reconstruct()',
    'Conditions': 'conditions argument',
    'Encoding_output': 'WaveNetAutoEncoder.encode',
    'Logits_e': 'WaveNetAutoEncoder.get_logits(inputs, encode(inputs))',
    'Out_e': 'WaveNetAutoEncoder.reconstruct',
    'Loss_e': 'WaveNetAutoEncoder.nll(inputs, encode(inputs))',
    'Encoding_input': 'encoding argument of get_logits / reconstruct_with_encoding',
    'Logits_d': 'WaveNetAutoEncoder.get_logits',
    'Out_d': 'WaveNetAutoEncoder.reconstruct_with_encoding',
    'Inputs_truth': 'inputs argument (teacher forcing)',
}
# what ParallelWaveNet.__init__ needs from the imported graph (model.py:326-341)
STUDENT_INPUT_MAP = ('inputs_truth_placeholder:0', 'conditions_placeholder:0', 'encoding_nodecoder_placeholder:0')
STUDENT_COLLECTIONS = ('Logits_d', 'Encoding_output', 'Inputs_e', 'Out_e', 'Out_d')


def read_meta(path):
    """-> dict(nodes={name: op}, collections={name: [tensor names]}) of a serialized MetaGraphDef."""
    with open(path, 'rb') as f:
        buf = memoryview(f.read())
    nodes, collections = {}, {}
    for num, wt, val in _fields(buf):
        if num == 2 and wt == 2:                                  # graph_def: GraphDef
            for n2, w2, v2 in _fields(val):
                if n2 == 1 and w2 == 2:                           # repeated NodeDef node
                    name = op = None
                    for n3, w3, v3 in _fields(v2):
                        if n3 == 1 and w3 == 2:
                            name = bytes(v3).decode()
                        elif n3 == 2 and w3 == 2:
                            op = bytes(v3).decode()
                    if name is not None:
                        nodes[name] = op
        elif num == 4 and wt == 2:                                # map<string, CollectionDef> collection_def entry
            key, items = None, []
            for n2, w2, v2 in _fields(val):
                if n2 == 1 and w2 == 2:
                    key = bytes(v2).decode()
                elif n2 == 2 and w2 == 2:                         # CollectionDef
                    for n3, w3, v3 in _fields(v2):
                        if n3 == 1 and w3 == 2:                   # NodeList node_list
                            items += [bytes(v4).decode() for n4, w4, v4 in _fields(v3) if n4 == 1 and w4 == 2]
                        elif n3 == 2 and w3 == 2:                 # BytesList (serialized VariableDef etc.): count only
                            items += [None for n4, w4, v4 in _fields(v3) if n4 == 1 and w4 == 2]
            if key is not None:
                collections[key] = items
    return dict(nodes=nodes, collections=collections)


def write_meta(path, nodes, collections):
    """nodes: {name: op}; collections: {name: [tensor names]} -> a MetaGraphDef file with that graph skeleton."""
    graph = b''.join(_ld(1, _ld(1, name.encode()) + _ld(2, op.encode())) for name, op in nodes.items())
    out = _ld(2, graph)
    for key, names in collections.items():
        node_list = b''.join(_ld(1, n.encode()) for n in names)
        out += _ld(4, _ld(1, key.encode()) + _ld(2, _ld(1, node_list)))
    with open(path, 'wb') as f:
        f.write(out)


def teacher_meta_skeleton(scope='WaveNetAutoEncoder'):
    """The placeholders and collections model.py:203-207 / 122-134 put into the teacher's graph.  Placeholder names and
    collection keys are the reference's literals (they are what the student's ``input_map`` and ``get_collection`` calls
    use); the tensor names listed INSIDE the collections are descriptive stand-ins, not TensorFlow's auto-generated op
    names (``.../conv1d_91/BiasAdd:0`` and the like), which nothing here depends on."""
    ph = ['inputs_placeholder', 'inputs_truth_placeholder', 'conditions_placeholder', 'encoding_nodecoder_placeholder']
    nodes = {'%s/%s' % (scope, p): 'Placeholder' for p in ph}
    t = lambda n: '%s/%s:0' % (scope, n)
    collections = {
        'Inputs_e': [t('inputs_placeholder')], 'Conditions': [t('conditions_placeholder')],
        'Encoding_output': [t('Encoder/AvgPool')], 'Logits_e': [t('Decoder/logits')], 'Out_e': [t('Decoder/out')],
        'Loss_e': ['loss:0'], 'Encoding_input': [t('encoding_nodecoder_placeholder')],
        'Logits_d': [t('Decoder_1/logits')], 'Out_d': [t('Decoder_1/out')], 'Inputs_truth': [t('inputs_truth_placeholder')],
    }
    return nodes, collections


def check_teacher_contract(meta, scope='WaveNetAutoEncoder'):
    """Raises what the student's constructor would raise on this graph (model.py:326-341)."""
    for key in STUDENT_INPUT_MAP:
        name = '%s/%s' % (scope, key)
        if name.split(':')[0] not in meta['nodes']:
            raise ValueError("input_map key %r is not a tensor of the imported teacher graph" % name)
    for c in STUDENT_COLLECTIONS:
        if not meta['collections'].get(c):
            raise IndexError("the teacher graph has no collection %r (model.py:122-134)" % c)
    return {c: meta['collections'][c][0] for c in STUDENT_COLLECTIONS}
