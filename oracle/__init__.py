"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the stable-ts word-timestamp hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
(or as the timed CPU arm), never on the B200 product path.

What is here
------------
``oracle.whisper_ref``  restatement of the third-party dependency that holds the arithmetic:
                        PyPI ``openai-whisper`` (pinned ``>=20230314,<=20250625`` by the reference's
                        setup.py:26-32; semantics of release 20250625).  It is NOT under
                        /root/reference and is not installed in this image, so its published
                        algorithm is restated and exposed with the same module/attribute names the
                        reference imports (stable_whisper/whisper_compatibility.py:58-76).
``oracle.stable_path``  restatement of the reference's own orchestration on the hot path
                        (stable_whisper/timing.py, decode.py, alignment.py:405-429,649-672).
``oracle/c``            plain-C restatement of the DTW and the width-7 median filter.
``oracle.silence``      numpy restatement of the non-VAD silence detection (SURVEY.md section 8f row 1:
                        stable_whisper/stabilization/nonvad.py, utils.py:43-111), pinned bit-exactly
                        against PyTorch's operators, the live reference functions and fixtures written
                        by the unmodified reference (oracle/make_golden_silence.py,
                        tests/test_oracle_silence.py).

Pinning status: the reference ships NO golden vectors for this path, so
"parity pinned by the reference's own tests" is impossible ("parity unpinned" in that sense).
The oracle is instead pinned against (1) the UNMODIFIED reference Python running here on top of
``oracle.whisper_ref`` registered as module ``whisper`` (tests/test_oracle_vs_reference.py and the
fixtures written by oracle/make_golden.py), and (2) independent ports of the same third-party
algorithms shipped in HF transformers (median filter, DTW, mel filter bank, Whisper forward).
"""
