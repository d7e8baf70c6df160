"""`from thinshelllab.engine.geometry import projection_query` -- the f_contact callback the reference's scripts pass
into Scene.time_step / Grad.transfer_grad (code/engine/geometry.py:223-229).  On the B200 path the candidate query is
part of the fused contact pipeline of libtsl (tsl_contact.cu), so the callable only marks that the CUDA query is
the one to run; calling it directly runs the query on the scene's engine."""


def projection_query(sys, debug=False):
    return sys.engine.contact_detect()
