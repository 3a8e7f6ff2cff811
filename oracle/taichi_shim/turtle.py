"""Stub: the reference's camera_isp.py:1 does `from turtle import color` (an editor auto-import);
tkinter is absent in this image.  TEST INFRASTRUCTURE ONLY."""
color = None
