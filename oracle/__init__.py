"""CPU oracle for the SpeechCLIP speech-image contrastive hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``speechclip_b200/`` or ``avssl/`` may
import this package: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker or the timed CPU baseline, never as the product path.

What it restates (plain fp32 torch on CPU, no CUDA):

* ``oracle.hubert``      fairseq HuBERT as driven by the reference wrapper
                         ``avssl/module/speech_encoder_plus.py:29-107,506-634``
* ``oracle.clip``        openai CLIP as driven by ``avssl/module/clip_official.py:200-264``
* ``oracle.speechclip``  weighted sum, key-padding mask, parallel branch, masked
                         InfoNCE, retrieval and the KWClip forward / loss
                         (``avssl/model/kwClip.py:1076-1108,1248-1297,1385-1478``,
                         ``avssl/module/{weighted_sum,losses,retrieval}.py``,
                         ``avssl/util/data_utils.py``)

Pinning status (see DESIGN.md "Oracle"):

* weighted sum, key-padding mask, branch encoder, masked InfoNCE, retrieval and
  the LR schedule are PINNED: ``tests/golden/make_golden.py`` imports the
  reference's own torch-only files from /root/reference and the committed
  fixtures in ``tests/golden/*.npz`` hold the reference's outputs.
* the HuBERT and CLIP towers are third-party code (fairseq @ b5a039c2, openai
  CLIP HEAD) that is ABSENT from /root/reference and from this image, and the
  reference's own tests hold no numeric vectors for them: **parity unpinned**
  against fairseq/openai themselves.  They are instead cross-checked against two
  independent in-container implementations of the same published architectures
  (``transformers.HubertModel`` / ``CLIPVisionModelWithProjection`` /
  ``CLIPTextModelWithProjection``) with weights mapped key-by-key, and those
  outputs are committed as fixtures too.
"""
