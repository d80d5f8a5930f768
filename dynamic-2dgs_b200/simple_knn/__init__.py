"""Drop-in for the reference's ``simple_knn`` package (submodules/simple-knn): ``from simple_knn._C import distCUDA2``
(scene/gaussian_model.py:20) resolves to the sm_100a kernel of libd2gs.so (d2gs_knn_mean_dist2)."""
