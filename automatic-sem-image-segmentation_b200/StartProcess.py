"""Drop-in for Releases/Version 1.2.0/StartProcess.py: same constants, same step functions, one spawned process per step
(:178-221; state travels between steps through the 1_WGAN / 2_CycleGAN / 3_UNet tree only).

Steps 1, 3, 4, 6a, 6b run networks on the sm_100a engine (WGAN-GP training, CycleGAN training / inference, MultiRes-UNet
training / inference).  Steps 0, 2 and 5 are host-side work around them (file handling, mask synthesis from the WGAN's
generator with simplex-noise clustering, classical post-processing), restated in HelperFunctions / WassersteinGAN /
Measurements.
"""
import os
from datetime import datetime

import multiprocessing as mp

from . import CycleGAN, HelperFunctions, UNet_Segmentation

ROOT_DIR = os.path.abspath("./")
INPUT_DIR_IMAGES = os.path.join(ROOT_DIR, "Input_Images")
OUTPUT_DIR_CYCLEGAN = os.path.join(ROOT_DIR, "Output_Masks_CycleGAN")
OUTPUT_DIR_UNET = os.path.join(ROOT_DIR, "Output_Masks_UNet")
TILE_SIZE_W = 384
TILE_SIZE_H = 384
RUN_INFERENCE_ON_WHOLE_IMAGE = True
USE_GPUS_NO = (0,)
ALLOW_MEMORY_GROWTH = True
INPUT_DIR_MASKS = os.path.join(ROOT_DIR, "Input_Masks")
WGAN_BATCH_SIZE = 64
WGAN_EPOCHS = 1000
MAX_PARTICLE_OVERLAP = 0.5
CYCLEGAN_BATCH_SIZE = 5
CYCLEGAN_EPOCHS = 50
CYCLEGAN_USE_SKIPS = False
CYCLEGAN_FILTERS = 64
UNET_BATCH_SIZE = 5
UNET_EPOCHS = 50
UNET_CONTRAST_OPTIMIZATION_RANGE = (0.5, 99.5)
UNET_FILTERS = 16
USE_DATALOADER = True
NUM_SIMULATED_MASKS = 1000
DARK_BACKGROUND = True
GAUSSIAN_BLUR_AMOUNT = 0.0


def start_step_0():
    print("Step0: Configuring Devices, Initializing Directories, and Preparing Images...")
    HelperFunctions.initialize_directories(root_dir=ROOT_DIR, output_dir_cyclegan=OUTPUT_DIR_CYCLEGAN, output_dir_unet=OUTPUT_DIR_UNET)
    HelperFunctions.prepare_images_cycle_gan(root_dir=ROOT_DIR, input_dir_images=INPUT_DIR_IMAGES, tile_size_w=TILE_SIZE_W,
                                             tile_size_h=TILE_SIZE_H, num_simulated_masks=NUM_SIMULATED_MASKS, dark_background=DARK_BACKGROUND)


def start_step_1():
    """StartProcess.py:61-68: train the WGAN-GP on the single-particle masks in Input_Masks."""
    from . import WassersteinGAN
    print('Step 1: Training WGAN...')
    wgan = WassersteinGAN.WGAN(root_dir=ROOT_DIR, allow_memory_growth=ALLOW_MEMORY_GROWTH, use_gpus_no=USE_GPUS_NO)
    wgan.batch_size = WGAN_BATCH_SIZE
    wgan.epochs = WGAN_EPOCHS
    wgan.n_z = 128
    return wgan.start_training()


def start_step_2():
    """StartProcess.py:71-87: simulate the fake masks (2_CycleGAN/data/trainB) with the WGAN trained in step 1."""
    from . import WassersteinGAN
    print('Step 2: Simulating fake masks...')
    num_masks = max(NUM_SIMULATED_MASKS, len(os.listdir(os.path.join(ROOT_DIR, '2_CycleGAN', 'data', 'trainA'))))
    w_gan = WassersteinGAN.WGAN(root_dir=ROOT_DIR, allow_memory_growth=ALLOW_MEMORY_GROWTH, use_gpus_no=USE_GPUS_NO)
    w_gan.n_z = 128
    w_gan.simulate_masks(no_of_images=num_masks, min_no_of_particles=100, max_no_of_particles=150, use_perlin_noise=True,
                         perlin_noise_threshold=0.5, perlin_noise_frequency=4, use_normal_distribution=True,
                         use_random_rotation='DISABLE', grid_type='DISABLE', max_overlap=MAX_PARTICLE_OVERLAP,
                         img_width=TILE_SIZE_W, img_height=TILE_SIZE_H)


def _cycle_gan():
    g = CycleGAN.CycleGAN(root_dir=ROOT_DIR, image_shape=(TILE_SIZE_H, TILE_SIZE_W, 1), allow_memory_growth=ALLOW_MEMORY_GROWTH,
                          use_gpus_no=USE_GPUS_NO)
    g.use_skip_connection = CYCLEGAN_USE_SKIPS
    g.filters = CYCLEGAN_FILTERS
    g.use_binary_crossentropy = False
    g.use_resize_convolution = False
    g.gaussian_noise_value = 0.0
    return g


def start_step_3():
    print("Step 3: Training CycleGAN...")
    g = _cycle_gan()
    g.batch_size = CYCLEGAN_BATCH_SIZE
    g.epochs = CYCLEGAN_EPOCHS
    g.use_data_loader = USE_DATALOADER
    g.label_smoothing_factor = 0.0
    g.lambda_identity_a = g.lambda_identity_b = 0.5
    g.start_training()


def start_step_4():
    print("Step 4: Generating fake training images and segmenting real images with CycleGAN...")
    g = _cycle_gan()
    g.run_inference(files=os.path.join(ROOT_DIR, "2_CycleGAN", "data", "trainB"),
                    output_directory=os.path.join(ROOT_DIR, "2_CycleGAN", "generate_images", "A"), source_domain="B",
                    tile_images=False, use_gpu=True)
    g.image_shape = (TILE_SIZE_W, TILE_SIZE_H)
    g.run_inference(files=INPUT_DIR_IMAGES, output_directory=os.path.join(ROOT_DIR, "2_CycleGAN", "generate_images", "B"),
                    source_domain="A", tile_images=not RUN_INFERENCE_ON_WHOLE_IMAGE, min_overlap=2, manage_overlap_mode=2, use_gpu=True)


def start_step_5():
    print("Step 5: Postprocessing CycleGAN Output images...")
    HelperFunctions.filter_gan_masks(img_path=os.path.join(ROOT_DIR, "2_CycleGAN", "generate_images", "A"),
                                     msk_path=os.path.join(ROOT_DIR, "2_CycleGAN", "data", "trainB"),
                                     out_path=os.path.join(ROOT_DIR, "2_CycleGAN", "generate_images", "Synthetic_Masks_Filtered"),
                                     gaussian_blur_amount=GAUSSIAN_BLUR_AMOUNT, do_watershed_and_four_connectivity=False,
                                     dark_background=DARK_BACKGROUND)
    HelperFunctions.filter_gan_masks(img_path=INPUT_DIR_IMAGES, msk_path=os.path.join(ROOT_DIR, "2_CycleGAN", "generate_images", "B"),
                                     out_path=OUTPUT_DIR_CYCLEGAN, do_watershed_and_four_connectivity=True, dark_background=DARK_BACKGROUND)


def _unet():
    u = UNet_Segmentation.UNet(root_dir=ROOT_DIR, image_dir=os.path.join(ROOT_DIR, "2_CycleGAN", "generate_images", "A"),
                               mask_dir=os.path.join(ROOT_DIR, "2_CycleGAN", "generate_images", "Synthetic_Masks_Filtered"),
                               allow_memory_growth=ALLOW_MEMORY_GROWTH, use_gpus_no=USE_GPUS_NO)
    u.use_dataloader = USE_DATALOADER
    u.filters = UNET_FILTERS
    u.contrast_optimization_range = UNET_CONTRAST_OPTIMIZATION_RANGE
    return u


def start_step_6a():
    print("Step 6.a: Train MultiRes UNet...")
    u = _unet()
    u.batch_size = UNET_BATCH_SIZE
    u.epochs = UNET_EPOCHS
    u.run_training()


def start_step_6b():
    print("Step 6.b: Segment real images with UNet")
    u = _unet()
    u.image_shape = (TILE_SIZE_W, TILE_SIZE_H)
    u.run_inference(files=INPUT_DIR_IMAGES, output_directory=OUTPUT_DIR_UNET, tile_images=not RUN_INFERENCE_ON_WHOLE_IMAGE,
                    threshold=-1, watershed_lines=True, min_distance=9, min_overlap=2, manage_overlap_mode=2, use_gpu=True)


def run_steps(steps):
    """One spawned process per step, joined before the next starts (reference :182-219; the parent does not look at exit
    codes there either, but a failed step is reported here)."""
    ctx = mp.get_context("spawn")
    for step in steps:
        p = ctx.Process(target=step)
        p.start()
        p.join()
        if p.exitcode != 0:
            print(f"{step.__name__} exited with code {p.exitcode}")


def main():
    """The reference's `__main__` (:178-221): all eight steps, one spawned process each."""
    print("Process started: " + str(datetime.now()))
    run_steps((start_step_0, start_step_1, start_step_2, start_step_3, start_step_4, start_step_5, start_step_6a, start_step_6b))
    print("Process finished: " + str(datetime.now()))


if __name__ == "__main__":
    main()
