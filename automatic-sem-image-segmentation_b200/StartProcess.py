"""Drop-in for the conv-stack steps of Releases/Version 1.2.0/StartProcess.py (steps 3, 4, 6a, 6b :89-175).

Steps 0-2 and 5 of the reference workflow (directory set-up, WGAN-GP mask synthesis, classical post-filtering) are
outside the accelerated path (SURVEY.md section 8 "out of scope" / "next"); run them with the reference and point
ROOT_DIR at the same tree.  Constants keep the reference's names and defaults.
"""
import os
from datetime import datetime

from . import CycleGAN, UNet_Segmentation

ROOT_DIR = os.path.abspath("./")
INPUT_DIR_IMAGES = os.path.join(ROOT_DIR, "Input_Images")
OUTPUT_DIR_CYCLEGAN = os.path.join(ROOT_DIR, "Output_Masks_CycleGAN")
OUTPUT_DIR_UNET = os.path.join(ROOT_DIR, "Output_Masks_UNet")
TILE_SIZE_W = 384
TILE_SIZE_H = 384
RUN_INFERENCE_ON_WHOLE_IMAGE = True
USE_GPUS_NO = (0,)
ALLOW_MEMORY_GROWTH = True
CYCLEGAN_BATCH_SIZE = 5
CYCLEGAN_EPOCHS = 50
CYCLEGAN_USE_SKIPS = False
CYCLEGAN_FILTERS = 64
UNET_BATCH_SIZE = 5
UNET_EPOCHS = 50
UNET_CONTRAST_OPTIMIZATION_RANGE = (0.5, 99.5)
UNET_FILTERS = 16
USE_DATALOADER = True


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


if __name__ == "__main__":
    print("Process started: " + str(datetime.now()))
    for step in (start_step_3, start_step_4, start_step_6a, start_step_6b):
        step()       # the reference isolates steps in processes to free TF GPU memory; torch frees buffers with the objects
    print("Process finished: " + str(datetime.now()))
