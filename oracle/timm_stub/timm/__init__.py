"""Minimal stand-in for timm==1.0.9 (pinned by the reference's environment.yaml:86).

TEST INFRASTRUCTURE ONLY. Lets the unmodified reference `model/HTR_VT.py` import in a
container where timm is not installed (reference import site: model_v1/model/HTR_VT.py:4).
"""
