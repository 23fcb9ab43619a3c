"""TEST INFRASTRUCTURE ONLY — CPU/fp32 oracle for the Flux-VAE HDR decode path.

Nothing under ``oracle/`` is imported by the product package
(``vae_decode_hdr_b200``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and
there only as the checker or as the timed CPU reference.

Parity pinning: the reference ships no tests or golden
vectors.  The HDR math restated in ``hdr_oracle.py`` is pinned against outputs
of the *unmodified* reference node (``/root/reference/hdr_vae_decode.py``)
executed in the build container; those outputs are committed under
``tests/golden/`` together with the generating script ``make_golden.py``.
The decoder arithmetic itself lives in ComfyUI (third party, un-vendored,
un-pinned: "provided by ComfyUI", reference requirements.txt:2); it is restated
in ``flux_decoder.py`` from the published BFL Flux.1 AE architecture and is
"parity unpinned" beyond "same module graph, same weights, fp32".
"""
