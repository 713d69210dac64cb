from nele_gan_b200.api import stoi  # noqa: F401
