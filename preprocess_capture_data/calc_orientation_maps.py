"""Drop-in module name of the reference (preprocess_capture_data/calc_orientation_maps.py), same command line."""
import argparse
import os

from monohair_b200.gabor import calc_confidences, calc_orients, generate_gabor_filters, main, rgb2gray  # noqa: F401

if __name__ == "__main__":
    parser = argparse.ArgumentParser(conflict_handler='resolve')
    root = 'data'
    parser.add_argument('--img_path', default=os.path.join(root, 'capture_images'), type=str)
    parser.add_argument('--orient_dir', default=os.path.join(root, 'orientation_maps'), type=str)
    parser.add_argument('--conf_dir', default=os.path.join(root, 'confidence_maps'), type=str)
    parser.add_argument('--mask_path', default=os.path.join(root, 'hair_mask'), type=str)
    parser.add_argument('--sigma_x', default=1.8, type=float)
    parser.add_argument('--sigma_y', default=2.4, type=float)
    parser.add_argument('--freq', default=0.23, type=float)
    parser.add_argument('--num_filters', default=180, type=int)
    main(parser.parse_args())
