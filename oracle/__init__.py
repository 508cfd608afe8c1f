"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's conversation-graph hot path
(sailist/emotion-recognition-in-conversation, track_mm/{cogmen,dgcn,mmgcn,dagerc}*).

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only as
the checker / the timed CPU baseline, never as the product path.  The product
package (``emotion-recognition-in-conversation_b200``) never imports ``oracle``.

PARITY STATUS -- "parity unpinned" for the PyG layers:
  * the reference has no tests, golden vectors or fixtures of its own;
  * the arithmetic of ``RGCNConv`` / ``TransformerConv`` / ``GraphConv`` used by
    COGMEN and DialogueGCN lives in an *unpinned, un-vendored* ``torch_geometric``
    (requirements.txt:12) and ``torch_scatter`` (models/rgcn.py:12), neither of
    which is installed here.  ``oracle/pyg_standin.py`` restates their published
    semantics in pure PyTorch; those stand-ins ARE the oracle for those layers.
  * everything that lives in the reference tree itself (edge_perms,
    batch_graphify, EdgeAtt, the vendored RGCNConv with edge_norm, the module
    forward functions) is pinned: ``oracle/make_golden.py`` imports the
    reference's own files from /root/reference (with the stand-ins pre-seeded in
    ``sys.modules``) and writes ``tests/golden/*.npz``; ``tests/test_oracle_*.py``
    check the restatement here against those fixtures.
"""
