"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the sparsify-clip loss hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU
baseline.  The product path (``sparsify_clip_b200``) never imports this package
and raises if its CUDA library is missing.

Contents
  closed_form.py  numpy fp64 restatement of the reference algorithm (values and
                  analytic gradients), each function citing the reference
                  file:line it follows.
  torch_port.py   the same op sequence the reference executes, in torch (CPU),
                  used for cpu_baseline timing ("kind": "port") and as a second
                  opinion for the fp64 closed forms.
  ref_loader.py   imports the UNMODIFIED reference from /root/reference (this
                  container only; absent on the GPU box) to pin the two
                  restatements above and to mint tests/golden/*.npz.
  make_golden.py  the script that generated tests/golden/.

Parity status: the reference ships no tests / golden vectors (SURVEY.md §4,
§8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in
this container through ref_loader.py; the resulting vectors are committed under
tests/golden/ together with make_golden.py.
"""
