"""`sonnet` (dm-sonnet 2.0.0) stand-in: `sonnet.src.{base,initializers,types,moving_averages}` as imported by the
reference's networks/vq_layers.py:6-9.  TEST INFRASTRUCTURE ONLY."""
