"""gan-des-midi-music-gen_b200: B200-native (sm_100a) MM-GAN / GAN-DES hot path.

Layout mirrors the reference's flat script directories:
    MMGAN_MIDI_DES.network_tests   Generator, BeatGenerator, Discriminator, DiscriminatorCNN, MultiModalGAN, ...
    MMGAN_MIDI_DES.datasets        generate_piano_roll, MaestroDataset*
    GAN_DES.SIMNN                  Generator, Discriminator, get_noise, weights_init
plus the fused pieces the reference has no name for (functional, optim, trainer).
"""
__version__ = "0.1.0"
