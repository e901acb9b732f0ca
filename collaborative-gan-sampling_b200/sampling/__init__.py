"""Drop-in mirror of the reference's ``sampling/`` directory (same module, class and method names).

Import either as a package (``from sampling.collaborator import Refiner``) or, like the reference does
(``sys.path.append('../sampling')`` at nsgan/GAN.py:15), put this directory on ``sys.path`` and
``from collaborator import Refiner``.
"""
